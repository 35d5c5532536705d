# Builds libsunet_b200.so (sm_100a only) in-tree next to the Python package.
NVCC ?= nvcc
PKG  := selectivenet_for_semantic_segmentation_binary_b200
SRC  := $(wildcard $(PKG)/csrc/*.cu)
HDR  := $(wildcard $(PKG)/csrc/*.h $(PKG)/csrc/*.cuh include/*.h)
OBJ  := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
LIB  := $(PKG)/libsunet_b200.so
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v

all: $(LIB)

build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) -shared -o $@ $(OBJ) -cudart static

# development probes (MMA-rate microbenchmark): NOT part of the product library
probes: scripts/probes/libsunet_probe.so
scripts/probes/libsunet_probe.so: scripts/probes/mma_probe.cu $(PKG)/csrc/common.cu $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ scripts/probes/mma_probe.cu $(PKG)/csrc/common.cu -cudart static

clean:
	rm -rf build $(LIB) scripts/probes/libsunet_probe.so
.PHONY: all clean probes
