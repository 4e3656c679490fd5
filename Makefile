# Developer entry points (reference: root Makefile lint/test targets + csrc/Makefile).
PY ?= python

.PHONY: build test test-gpu lint sanitize bench clean

build:            ## compile csrc/*.cu (sm_100a) and the C++ runtime in-tree
	$(PY) -m easyparallellibrary_b200.build

test:             ## CPU tiers (gloo, multi-process)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## kernel numerics + multi-GPU parity (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

lint:             ## syntax / import check of the package, -Wall build of the host runtime
	$(PY) -m compileall -q easyparallellibrary_b200 tests examples tools bench.py
	$(PY) tools/check_globals.py          # every global name and every ctypes symbol resolves (GPU-only branches included)
	g++ -O2 -std=c++17 -fPIC -Wall -Werror=return-type -fsyntax-only -I/usr/local/cuda/include easyparallellibrary_b200/csrc/runtime.cpp

sanitize:         ## memory / race / sync checks of the kernels (needs a B200; slow: kernels run ~50x slower)
	compute-sanitizer --tool memcheck  $(PY) -m pytest tests/test_kernels_gpu.py -x -q -k "adamw or layernorm or xent or rope"
	compute-sanitizer --tool racecheck $(PY) -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm or xent"
	compute-sanitizer --tool synccheck $(PY) -m pytest tests/test_kernels_gpu.py -x -q -k "gemm_layouts or flash_attention"

bench:
	$(PY) bench.py --gpus 1 --steps 10 --warmup 3

clean:
	rm -rf easyparallellibrary_b200/lib build *.egg-info
