# Builds libforagax_b200.so (sm_100a only) in-tree and the C oracle used by tests/bench.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Iinclude
CSRC      := foragax_b200/csrc
BUILD     := build/obj
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(SRCS))
LIB       := foragax_b200/libforagax_b200.so

# geometry kernels must not contract mul+add into FMA (discrete orientation predicates)
NOFMA     := raycast

all: $(LIB) oracle

$(BUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/foragax_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) $(if $(filter $*,$(NOFMA)),-fmad=false,) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared --cudart static -o $@ $(OBJS)

oracle: oracle/_build/libfgx_oracle.so

oracle/_build/libfgx_oracle.so: oracle/c/fgx_oracle.c
	@mkdir -p oracle/_build
	gcc -O3 -mavx2 -mno-fma -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math -o $@ $< -lm

# compile-only check of the XLA-FFI shim against a stand-in for jaxlib's xla/ffi/api/ffi.h (jaxlib is absent from
# this image): every handler's parameter list is static_assert-ed against its binding and every launcher call
# against include/foragax_b200.h
xla-ffi-check:
	g++ -std=c++17 -fsyntax-only -I$(CSRC)/xla_ffi/stub -Iinclude -I/usr/local/cuda/include $(CSRC)/xla_ffi/xla_ffi_shim.cc

ptxas-info:
	@for f in $(SRCS); do echo "== $$f"; $(NVCC) $(NVFLAGS) -Xptxas -v -c $$f -o /dev/null 2>&1 | grep -E "Compiling|registers|spill" ; done

clean:
	rm -rf build $(LIB) oracle/_build

.PHONY: all oracle clean ptxas-info xla-ffi-check
