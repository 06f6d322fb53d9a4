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

# instrumented builds of the persistent tcgen05 kernels (cycle accounting; read by tools/trace_gconv2.py and
# tools/dbg_wgrad.py with G2LIB= / WGLIB=), and the hardware probes
TRACEFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude -shared
trace:
	$(NVCC) $(TRACEFLAGS) -DHDMOE_G2_TRACE -o tools/libg2trace.so $(PKG)/csrc/gconv2.cu $(PKG)/csrc/core.cu -lcudart
	$(NVCC) $(TRACEFLAGS) -DHDMOE_WG_TRACE -o tools/libwg2trace.so $(PKG)/csrc/gwgrad2.cu $(PKG)/csrc/core.cu -lcudart
probes:
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -o tools/umma_tpair_probe tools/umma_tpair_probe.cu

clean:
	rm -rf build $(LIB)
.PHONY: all clean trace probes
