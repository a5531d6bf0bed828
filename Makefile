# Native build of the C-ABI library and its C++ host example (no Python needed).
# `python -c "import __graft_entry__ as g; g.build()"` does the same through prcv2025reid_b200/build.py.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr
PKG       := prcv2025reid_b200
SRC       := normalize pos_index rank sdm sdm_tc sim_gemm retrieve_fused api
OBJ       := $(SRC:%=$(PKG)/build/%.o)
HDR       := $(wildcard $(PKG)/csrc/*.cuh) include/reid_b200.h

.PHONY: all lib example clean
all: lib example
lib: $(PKG)/libreid_b200.so
example: examples/abi_host

$(PKG)/build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p $(PKG)/build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(PKG)/libreid_b200.so: $(OBJ)
	$(NVCC) -shared -o $@ $(OBJ)

examples/abi_host: examples/abi_host.cpp include/reid_b200.h $(PKG)/libreid_b200.so
	$(NVCC) -O2 -std=c++17 -Wno-deprecated-gpu-targets -Iinclude $< -o $@ -L$(PKG) -lreid_b200 \
	    -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)'

clean:
	rm -rf $(PKG)/build $(PKG)/libreid_b200.so examples/abi_host
