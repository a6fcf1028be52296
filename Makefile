# Builds libpsagpu.so (hand-written CUDA for sm_100a + C ABI) and the CPU oracle used by the tests.
NVCC ?= nvcc
PKG := pseudospectra.jl_b200
CSRC := $(PKG)/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
LIB := $(PKG)/libpsagpu.so

all: $(LIB) oracle

$(LIB): $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.h) include/psa_gpu.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/psa_api.cu -ldl

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
