PY ?= python
.PHONY: build test test-gpu bench sanitize asan clean
build:
	$(PY) -m infomesh_b200.build
test:
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu:
	$(PY) -m pytest tests -x -q -m gpu
bench:
	$(PY) bench.py --gpus 1 --steps 10 --warmup 3
sanitize:
	scripts/sanitize.sh memcheck && scripts/sanitize.sh racecheck && scripts/sanitize.sh synccheck
asan:   # C++ host runtime under AddressSanitizer + UBSan (tokenizer / posting builder / SimHash)
	g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -shared -fPIC -o /tmp/libtextproc_asan.so infomesh_b200/csrc/host/textproc.cpp
	@echo "LD_PRELOAD=$$(gcc -print-file-name=libasan.so) INFOMESH_TEXTPROC_LIB=/tmp/libtextproc_asan.so $(PY) -m pytest tests/test_native_cpu.py -q"
clean:
	rm -rf infomesh_b200/_native build *.egg-info
