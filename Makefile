# Builds the product: ising_gpu_b200/libising_b200.so (C ABI, include/ising_b200.h) and the
# cuIsing command-line host (reference CLI surface).  sm_100a only.
NVCC    ?= /usr/local/cuda/bin/nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
CSRC    := ising_gpu_b200/csrc
LIB     := ising_gpu_b200/libising_b200.so
CLI     := ising_gpu_b200/cuIsing
CLI_TC  := ising_gpu_b200/ising_tensorcore

all: $(LIB) $(CLI) $(CLI_TC) oracle

$(LIB): $(CSRC)/kernels.cu $(CSRC)/ising_api.cu $(CSRC)/tc_variant.cu $(CSRC)/kernels.cuh $(CSRC)/philox_map.h $(CSRC)/host_math.h include/ising_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/kernels.cu $(CSRC)/ising_api.cu $(CSRC)/tc_variant.cu

$(CLI): $(CSRC)/cuising_main.cpp $(LIB) include/ising_b200.h
	$(NVCC) -O2 -std=c++17 -Xcompiler -Wall -o $@ $(CSRC)/cuising_main.cpp -Iinclude -Lising_gpu_b200 -lising_b200 -Xlinker -rpath -Xlinker '$$ORIGIN'

$(CLI_TC): $(CSRC)/ising_tensorcore_main.cpp $(LIB) include/ising_b200.h
	$(NVCC) -O2 -std=c++17 -Xcompiler -Wall -o $@ $(CSRC)/ising_tensorcore_main.cpp -Iinclude -Lising_gpu_b200 -lising_b200 -Xlinker -rpath -Xlinker '$$ORIGIN'

oracle:
	$(MAKE) -C oracle

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $(CSRC)/kernels.cu -o /dev/null

clean:
	rm -f $(LIB) $(CLI) $(CLI_TC)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean ptxas-info
