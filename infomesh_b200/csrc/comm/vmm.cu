// Peer-memory backend, second generation (SURVEY C-0 / §5.8): a symmetric heap built on the CUDA VMM API
// (cuMemCreate / export as POSIX fd / import / cuMemMap) plus an NVLS multicast object over the same physical memory
// (cuMulticastCreate / AddDevice / BindMem), and the kernels that use the multicast address:
//
//   * mc_allgather_kernel   one `multimem.st` per 16 bytes lands a rank's block in EVERY GPU's receive slot (the switch
//                           replicates the store), one `multimem.red` bumps every GPU's arrival counter -- versus
//                           world stores + world reds over unicast P2P (symm.cu).  Used for the query-embedding broadcast
//                           of the de-replicated encoder and the logits gather.
//   * mc_reduce_scatter_kernel   `multimem.ld_reduce.add` (bf16x2, fp32 accumulate): the switch pulls the same address
//                           from all GPUs and returns the sum -- the GEMM -> reduce-scatter consumer without P partial
//                           buffers or P reads; optionally writes the reduced rows back through `multimem.st`
//                           (all-reduce).
//
// The driver entry points are resolved at run time (the library must link on the GPU-less build box).  File descriptors
// are passed between the rank processes by the Python side (SCM_RIGHTS over a Unix socket, parallel/vmm.py).
// Replaces the scatter / gather the reference does over WAN streams (infomesh/p2p/routing.py:193-267).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <unistd.h>

#include <mutex>

#include "../common/host.h"
#include "../common/ptx.cuh"

namespace im {

struct DriverFns {
  CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*memRelease)(CUmemGenericAllocationHandle);
  CUresult (*memExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*memImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*memAddressFree)(CUdeviceptr, size_t);
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*memUnmap)(CUdeviceptr, size_t);
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*memGetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*mcCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*mcAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*mcBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*mcGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*deviceGet)(CUdevice*, int);
  CUresult (*deviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  bool ok = false;
};

static DriverFns& drv() {
  static DriverFns f;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char* name, void** out) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, out, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess &&
             *out != nullptr;
    };
    bool ok = true;
    ok &= get("cuMemCreate", reinterpret_cast<void**>(&f.memCreate));
    ok &= get("cuMemRelease", reinterpret_cast<void**>(&f.memRelease));
    ok &= get("cuMemExportToShareableHandle", reinterpret_cast<void**>(&f.memExport));
    ok &= get("cuMemImportFromShareableHandle", reinterpret_cast<void**>(&f.memImport));
    ok &= get("cuMemAddressReserve", reinterpret_cast<void**>(&f.memAddressReserve));
    ok &= get("cuMemAddressFree", reinterpret_cast<void**>(&f.memAddressFree));
    ok &= get("cuMemMap", reinterpret_cast<void**>(&f.memMap));
    ok &= get("cuMemUnmap", reinterpret_cast<void**>(&f.memUnmap));
    ok &= get("cuMemSetAccess", reinterpret_cast<void**>(&f.memSetAccess));
    ok &= get("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&f.memGetGranularity));
    ok &= get("cuDeviceGet", reinterpret_cast<void**>(&f.deviceGet));
    ok &= get("cuDeviceGetAttribute", reinterpret_cast<void**>(&f.deviceGetAttribute));
    // multicast entry points may be absent on old drivers: VMM still works without them
    get("cuMulticastCreate", reinterpret_cast<void**>(&f.mcCreate));
    get("cuMulticastAddDevice", reinterpret_cast<void**>(&f.mcAddDevice));
    get("cuMulticastBindMem", reinterpret_cast<void**>(&f.mcBindMem));
    get("cuMulticastGetGranularity", reinterpret_cast<void**>(&f.mcGetGranularity));
    f.ok = ok;
  });
  return f;
}

static int cu_fail(const char* what, CUresult r) {
  char msg[96];
  snprintf(msg, sizeof msg, "CUresult=%d", static_cast<int>(r));
  return set_error(what, msg);
}
#define IM_CU_OK(expr)                        \
  do {                                        \
    CUresult _r = (expr);                     \
    if (_r != CUDA_SUCCESS) return cu_fail(#expr, _r); \
  } while (0)

static CUmemAllocationProp alloc_prop(int dev) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

// ---------------------------------------------------------------------------------------------------------------
// multimem kernels
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mc_flag_wait(const uint32_t* flag, uint32_t target, const char* what) {
  uint32_t spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys(flag) - target) < 0) {
    if (++spins > IM_WAIT_LIMIT) {
      printf("[infomesh_b200] %s timeout (have %u want %u)\n", what, ld_acquire_sys(flag), target);
      __trap();
    }
    __nanosleep(20);
  }
}
__device__ __forceinline__ void mc_channel_advance(uint32_t* state) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 1, 1u) == gridDim.x - 1u) {
      state[1] = 0u;
      state[0] += 1u;
      __threadfence();
    }
  }
}

// All-gather through the switch.  mc_buf: multicast VA of the receive area [2][world][bytes]; local_buf: this rank's
// unicast VA of the same area; mc_flag / local_flag: [world] arrival counters (multicast / unicast view).
__global__ void __launch_bounds__(256)
mc_allgather_kernel(const uint8_t* __restrict__ src, size_t bytes, uint8_t* mc_buf, const uint8_t* local_buf, uint32_t* mc_flag,
                    const uint32_t* local_flag, uint32_t* state, int world, int rank, uint8_t* __restrict__ out) {
  const uint32_t step = *reinterpret_cast<volatile uint32_t*>(state);
  const size_t par_off = static_cast<size_t>(step & 1u) * world * bytes;
  const size_t n16 = bytes / 16;
  const size_t per_cta = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per_cta, hi = min(n16, lo + per_cta);
  uint8_t* dst = mc_buf + par_off + static_cast<size_t>(rank) * bytes;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) multimem_st_v4(dst + i * 16, s4[i]);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) multimem_red_add_release(mc_flag + rank, 1u);     // every GPU's counter[rank] += 1, one instruction
  const uint32_t target = (step + 1u) * gridDim.x;
  if (threadIdx.x < static_cast<unsigned>(world)) mc_flag_wait(local_flag + threadIdx.x, target, "mc_allgather");
  __syncthreads();
  if (out != nullptr) {
    const uint8_t* mine = local_buf + par_off;
    for (int p = 0; p < world; ++p) {
      const uint4* g4 = reinterpret_cast<const uint4*>(mine + static_cast<size_t>(p) * bytes);
      uint4* d4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(p) * bytes);
      for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) d4[i] = g4[i];
    }
  }
  mc_channel_advance(state);
}

// Reduce-scatter (all_reduce = 0) or all-reduce (= 1) of a symmetric bf16 / fp32 buffer through the switch.
// Every rank has written its partial into the same symmetric offset (unicast).  Barrier protocol (two counter bumps per
// rank and use, so late-wave CTAs never deadlock against peers):
//   entry: CTA 0 arrives once (the kernel is stream-ordered after the producer, so a running CTA means the partial is
//          complete); every CTA waits for all ranks' entry arrivals, then reduces its slice with multimem.ld_reduce
//          (4 independent 16-byte reductions in flight per thread);
//   exit:  the last CTA out arrives and waits for every rank's exit arrival before the kernel may complete, so a fast
//          rank cannot overwrite its partial (next use) while a slower peer is still pulling it through the switch.
__global__ void __launch_bounds__(256)
mc_reduce_kernel(const uint8_t* mc_in, uint8_t* mc_out, uint8_t* __restrict__ out, size_t n16_total, int is_f32, uint32_t* mc_flag,
                 const uint32_t* local_flag, uint32_t* state, int world, int rank, int all_reduce) {
  const uint32_t step = *reinterpret_cast<volatile uint32_t*>(state);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    multimem_red_add_release(mc_flag + rank, 1u);
  }
  if (threadIdx.x < static_cast<unsigned>(world)) mc_flag_wait(local_flag + threadIdx.x, step * 2u + 1u, "mc_reduce(enter)");
  __syncthreads();
  const size_t per_rank = (n16_total + world - 1) / world;
  const size_t r_lo = min(n16_total, static_cast<size_t>(rank) * per_rank), r_hi = min(n16_total, r_lo + per_rank);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i0 = r_lo + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < r_hi; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = i0 + j * stride;
      if (i < r_hi) {
        if (is_f32) {
          const float4 f = multimem_ld_reduce_f32x4(mc_in + i * 16);
          v[j] = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
        } else {
          v[j] = multimem_ld_reduce_bf16x8(mc_in + i * 16);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = i0 + j * stride;
      if (i < r_hi) {
        if (all_reduce) multimem_st_v4(mc_out + i * 16, v[j]);
        else reinterpret_cast<uint4*>(out)[i - r_lo] = v[j];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(state + 1, 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) multimem_red_add_release(mc_flag + rank, 1u);
    if (threadIdx.x < static_cast<unsigned>(world)) mc_flag_wait(local_flag + threadIdx.x, step * 2u + 2u, "mc_reduce(exit)");
    __syncthreads();
    if (threadIdx.x == 0) {
      state[1] = 0u;
      state[0] = step + 1u;
      __threadfence();
    }
  }
}

}  // namespace im

using namespace im;

// bit0 VMM, bit1 POSIX-fd handles, bit2 multicast (NVLS)
IM_API int im_vmm_supported(int dev) {
  DriverFns& d = drv();
  if (!d.ok) return 0;
  CUdevice cd;
  if (d.deviceGet(&cd, dev) != CUDA_SUCCESS) return 0;
  int vmm = 0, fd = 0, mc = 0;
  d.deviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cd);
  d.deviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cd);
  d.deviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd);
  if (d.mcCreate == nullptr) mc = 0;
  return (vmm ? 1 : 0) | (fd ? 2 : 0) | (mc ? 4 : 0);
}

// Round `bytes` up to the granularity both a device allocation and a multicast object of `world` devices accept.
IM_API long long im_vmm_round_size(long long bytes, int dev, int world) {
  DriverFns& d = drv();
  if (!d.ok) return -1;
  CUmemAllocationProp p = alloc_prop(dev);
  size_t g = 0;
  if (d.memGetGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || g == 0) g = 2u << 20;
  if (d.mcGetGranularity != nullptr && world > 1) {
    CUmulticastObjectProp mp = {};
    mp.numDevices = static_cast<unsigned>(world);
    mp.size = static_cast<size_t>(bytes);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (d.mcGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
  }
  return static_cast<long long>((static_cast<size_t>(bytes) + g - 1) / g * g);
}

IM_API int im_vmm_create(long long size, int dev, unsigned long long* handle_out, int* fd_out) {
  DriverFns& d = drv();
  if (!d.ok) return set_error("im_vmm_create", "CUDA VMM driver entry points unavailable");
  CUmemAllocationProp p = alloc_prop(dev);
  CUmemGenericAllocationHandle h;
  IM_CU_OK(d.memCreate(&h, static_cast<size_t>(size), &p, 0));
  int fd = -1;
  CUresult r = d.memExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    d.memRelease(h);
    return cu_fail("cuMemExportToShareableHandle", r);
  }
  *handle_out = static_cast<unsigned long long>(h);
  *fd_out = fd;
  return 0;
}

IM_API int im_vmm_import(int fd, unsigned long long* handle_out) {
  DriverFns& d = drv();
  if (!d.ok) return set_error("im_vmm_import", "CUDA VMM driver entry points unavailable");
  CUmemGenericAllocationHandle h;
  IM_CU_OK(d.memImport(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  *handle_out = static_cast<unsigned long long>(h);
  return 0;
}

// Reserve a VA range, map `handle` (a memory allocation OR a multicast object) and grant `dev` read/write access.
IM_API int im_vmm_map(unsigned long long handle, long long size, int dev, void** ptr_out) {
  DriverFns& d = drv();
  if (!d.ok) return set_error("im_vmm_map", "CUDA VMM driver entry points unavailable");
  CUdeviceptr va = 0;
  IM_CU_OK(d.memAddressReserve(&va, static_cast<size_t>(size), 0, 0, 0));
  CUresult r = d.memMap(va, static_cast<size_t>(size), 0, static_cast<CUmemGenericAllocationHandle>(handle), 0);
  if (r != CUDA_SUCCESS) {
    d.memAddressFree(va, static_cast<size_t>(size));
    return cu_fail("cuMemMap", r);
  }
  CUmemAccessDesc a = {};
  a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a.location.id = dev;
  a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.memSetAccess(va, static_cast<size_t>(size), &a, 1);
  if (r != CUDA_SUCCESS) {
    d.memUnmap(va, static_cast<size_t>(size));
    d.memAddressFree(va, static_cast<size_t>(size));
    return cu_fail("cuMemSetAccess", r);
  }
  *ptr_out = reinterpret_cast<void*>(va);
  return 0;
}

IM_API int im_vmm_unmap(void* ptr, long long size) {
  DriverFns& d = drv();
  if (!d.ok || ptr == nullptr) return 0;
  d.memUnmap(reinterpret_cast<CUdeviceptr>(ptr), static_cast<size_t>(size));
  d.memAddressFree(reinterpret_cast<CUdeviceptr>(ptr), static_cast<size_t>(size));
  return 0;
}

IM_API int im_vmm_release(unsigned long long handle) {
  DriverFns& d = drv();
  if (d.ok && handle != 0) d.memRelease(static_cast<CUmemGenericAllocationHandle>(handle));
  return 0;
}

IM_API int im_vmm_close_fd(int fd) {
  if (fd >= 0) close(fd);
  return 0;
}

// ---- NVLS multicast object ----
IM_API int im_mc_create(long long size, int world, unsigned long long* handle_out, int* fd_out) {
  DriverFns& d = drv();
  if (!d.ok || d.mcCreate == nullptr) return set_error("im_mc_create", "cuMulticast* entry points unavailable");
  CUmulticastObjectProp mp = {};
  mp.numDevices = static_cast<unsigned>(world);
  mp.size = static_cast<size_t>(size);
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  IM_CU_OK(d.mcCreate(&h, &mp));
  int fd = -1;
  CUresult r = d.memExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    d.memRelease(h);
    return cu_fail("cuMemExportToShareableHandle(multicast)", r);
  }
  *handle_out = static_cast<unsigned long long>(h);
  *fd_out = fd;
  return 0;
}

IM_API int im_mc_add_device(unsigned long long mc, int dev) {
  DriverFns& d = drv();
  if (!d.ok || d.mcAddDevice == nullptr) return set_error("im_mc_add_device", "cuMulticast* entry points unavailable");
  CUdevice cd;
  IM_CU_OK(d.deviceGet(&cd, dev));
  IM_CU_OK(d.mcAddDevice(static_cast<CUmemGenericAllocationHandle>(mc), cd));
  return 0;
}

// Bind this device's physical allocation at offset 0 of the multicast object (every device must have been added first).
IM_API int im_mc_bind(unsigned long long mc, unsigned long long mem, long long size) {
  DriverFns& d = drv();
  if (!d.ok || d.mcBindMem == nullptr) return set_error("im_mc_bind", "cuMulticast* entry points unavailable");
  IM_CU_OK(d.mcBindMem(static_cast<CUmemGenericAllocationHandle>(mc), 0, static_cast<CUmemGenericAllocationHandle>(mem), 0,
                       static_cast<size_t>(size), 0));
  return 0;
}

// ---- device primitives on the multicast mapping ----
IM_API int im_mc_allgather(const void* src, size_t bytes, void* mc_buf, const void* local_buf, uint32_t* mc_flag,
                           const uint32_t* local_flag, uint32_t* state, int world, int rank, void* out, int ctas, void* stream) {
  if (bytes == 0 || (bytes % 16) != 0) return set_error("im_mc_allgather", "block size must be a non-zero multiple of 16 bytes");
  const size_t n16 = bytes / 16;
  int grid = ctas > 0 ? ctas : static_cast<int>((n16 + 1023) / 1024);
  grid = grid < 1 ? 1 : (grid > 32 ? 32 : grid);
  mc_allgather_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), bytes, static_cast<uint8_t*>(mc_buf), static_cast<const uint8_t*>(local_buf), mc_flag,
      local_flag, state, world, rank, static_cast<uint8_t*>(out));
  IM_LAUNCH_OK("mc_allgather_kernel");
  return grid;
}

IM_API int im_mc_reduce(const void* mc_in, void* mc_out, void* out, size_t bytes, int is_f32, uint32_t* mc_flag,
                        const uint32_t* local_flag, uint32_t* state, int world, int rank, int all_reduce, int ctas, void* stream) {
  if (bytes == 0 || (bytes % 16) != 0) return set_error("im_mc_reduce", "buffer size must be a non-zero multiple of 16 bytes");
  const size_t n16 = bytes / 16;
  int grid = ctas > 0 ? ctas : static_cast<int>((n16 / world + 1023) / 1024);
  grid = grid < 1 ? 1 : (grid > 2 * sm_count() ? 2 * sm_count() : grid);
  mc_reduce_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(mc_in), static_cast<uint8_t*>(mc_out), static_cast<uint8_t*>(out), n16, is_f32, mc_flag, local_flag,
      state, world, rank, all_reduce);
  IM_LAUNCH_OK("mc_reduce_kernel");
  return grid;
}
