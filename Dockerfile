# B200 node image: CUDA 12.9 toolchain to build the sm_100a kernels in-tree, PyTorch runtime, non-root user.
FROM nvidia/cuda:12.9.0-devel-ubuntu24.04 AS build
RUN apt-get update && apt-get install -y --no-install-recommends python3 python3-venv python3-pip g++ libzstd1 && rm -rf /var/lib/apt/lists/*
WORKDIR /app
RUN python3 -m venv /app/.venv
ENV PATH="/app/.venv/bin:$PATH"
COPY pyproject.toml README.md ./
COPY infomesh_b200/ infomesh_b200/
RUN pip install --no-cache-dir ".[all]" && python -m infomesh_b200.build

FROM nvidia/cuda:12.9.0-runtime-ubuntu24.04
RUN apt-get update && apt-get install -y --no-install-recommends python3 libzstd1 curl && rm -rf /var/lib/apt/lists/* \
    && groupadd --gid 1000 infomesh && useradd --uid 1000 --gid infomesh --create-home infomesh
COPY --from=build /app /app
ENV PATH="/app/.venv/bin:$PATH" PYTHONUNBUFFERED=1 INFOMESH_NODE_DATA_DIR=/data INFOMESH_GPU_ENABLED=true
RUN mkdir -p /data && chown infomesh:infomesh /data
USER infomesh
# 4001 P2P · 8080 admin API · 8081 MCP streamable HTTP
EXPOSE 4001 8080 8081
VOLUME /data
HEALTHCHECK --interval=30s --timeout=5s --start-period=20s CMD curl -fs http://127.0.0.1:8080/health || exit 1
ENTRYPOINT ["python3", "-m", "infomesh_b200"]
CMD ["_serve"]
