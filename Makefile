# Builds, in-tree:
#   oswald_b200/liboswald_cuda.so   the product: CUDA kernels (sm_100a) + C ABI (include/oswald_cuda.h)
#   oswald_b200/oswald              the command-line tool (host C, reference CLI) linked against it
#   tools/osw_synth, tools/libosw_synth.so   synthetic data generator
#   oracle/...                      the parity checker (test infrastructure; see oracle/Makefile)
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCC   := $(firstword $(wildcard /usr/bin/gcc) gcc)
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -Iinclude
CUSRC    := $(wildcard oswald_b200/csrc/cuda/*.cu)
CUOBJ    := $(patsubst oswald_b200/csrc/cuda/%.cu,build/%.o,$(CUSRC))
HOSTLIB  := oswald_b200/csrc/host/dbformat.c oswald_b200/csrc/host/submat.c
CLISRC   := $(filter-out $(HOSTLIB),$(wildcard oswald_b200/csrc/host/*.c))

all: lib cli tools oracle

lib: oswald_b200/liboswald_cuda.so oswald_b200/build_tag.txt
cli: oswald_b200/oswald
tools: tools/osw_synth tools/libosw_synth.so

build/%.o: oswald_b200/csrc/cuda/%.cu oswald_b200/csrc/cuda/osw_internal.h oswald_b200/csrc/cuda/sw_t16.h oswald_b200/csrc/cuda/sw_u16_kernel.cuh include/oswald_cuda.h oswald_b200/csrc/host/dbformat.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

build/%.o: oswald_b200/csrc/host/%.c $(wildcard oswald_b200/csrc/host/*.h) oswald_b200/csrc/host/submat_tri.inc
	@mkdir -p build
	$(HOSTCC) -O2 -fPIC -std=c11 -Wall -fopenmp -Iinclude -c $< -o $@

oswald_b200/liboswald_cuda.so: $(CUOBJ) build/dbformat.o build/submat.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart_static -lpthread -ldl -lrt -lgomp

# hash of the kernel, planner and layout sources the library was built from (not of the host-side
# orchestration in api.cu or the calibration kernels): bench.py reports it as `build` and matches it
# against the tag stored with profiles/ncu_traffic.json
TAGSRC   := $(sort $(wildcard oswald_b200/csrc/cuda/sw_*.cu oswald_b200/csrc/cuda/*.cuh oswald_b200/csrc/cuda/*.h)) \
            oswald_b200/csrc/cuda/topr.cu oswald_b200/csrc/cuda/plan.cu oswald_b200/csrc/host/dbformat.c oswald_b200/csrc/host/dbformat.h
oswald_b200/build_tag.txt: oswald_b200/liboswald_cuda.so
	cat $(TAGSRC) | sha256sum | cut -c1-12 > $@

oswald_b200/oswald: $(CLISRC) oswald_b200/liboswald_cuda.so
	$(HOSTCC) -O2 -std=gnu11 -Wall -fopenmp -Iinclude -Ioswald_b200/csrc/host -o $@ $(CLISRC) \
	    -Loswald_b200 -loswald_cuda -Wl,-rpath,'$$ORIGIN' -lm

tools/osw_synth: tools/osw_synth.c
	$(HOSTCC) -O2 -fopenmp -o $@ $< -lm
tools/libosw_synth.so: tools/osw_synth.c
	$(HOSTCC) -O2 -fopenmp -fPIC -shared -DOSW_SYNTH_NO_MAIN -o $@ $< -lm

# experiment: dependent-issue latency of the DPX instructions (not part of `all`)
tools/dpx_latency: tools/dpx_latency.cu
	$(NVCC) $(ARCH) -O3 -o $@ $<

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build oswald_b200/liboswald_cuda.so oswald_b200/build_tag.txt oswald_b200/oswald tools/osw_synth tools/libosw_synth.so tools/dpx_latency
	$(MAKE) -C oracle clean
.PHONY: all lib cli tools oracle clean
