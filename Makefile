# Builds libhdmoe_b200.so (hand-written sm_100a kernels behind a C ABI) and the oracle helpers.
NVCC      ?= nvcc
PKG       := heterogeneous-moe-for-diffusion-models_b200
SRC       := $(wildcard $(PKG)/csrc/*.cu)
OBJ       := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
LIB       := $(PKG)/lib/libhdmoe_b200.so
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
             -Xptxas -v --expt-relaxed-constexpr -Iinclude

all: $(LIB)

build/%.o: $(PKG)/csrc/%.cu $(wildcard $(PKG)/csrc/*.cuh) $(wildcard include/*.h)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p $(PKG)/lib
	$(NVCC) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean
