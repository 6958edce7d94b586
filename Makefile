# Builds vv_dsp_b200/lib/libvvdsp_b200.so without Python: C99 host sources with gcc, one CUDA translation unit per kernel
# family with nvcc for sm_100a only (the same commands as `python -m vv_dsp_b200.build`, which __graft_entry__.build() uses).
#   make -j            the library
#   make oracle        the CPU checker (test infrastructure; `ref` needs /root/reference)
#   make c-callers     the reference-style C programs of tests/c against the library
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
PKG  := vv_dsp_b200
OBJ  := $(PKG)/lib/obj
LIB  := $(PKG)/lib/libvvdsp_b200.so
HOST := $(wildcard $(PKG)/csrc/host/*.c)
CUDA := $(wildcard $(PKG)/csrc/cuda/*.cu)
HDRS := $(shell find include -name '*.h') $(wildcard $(PKG)/csrc/cuda/*.cuh) $(wildcard $(PKG)/csrc/host/*.h)
OBJS := $(patsubst $(PKG)/csrc/host/%.c,$(OBJ)/%.c.o,$(HOST)) $(patsubst $(PKG)/csrc/cuda/%.cu,$(OBJ)/%.cu.o,$(CUDA))

all: $(LIB)

$(OBJ)/%.c.o: $(PKG)/csrc/host/%.c $(HDRS)
	@mkdir -p $(OBJ)
	gcc -std=c99 -O2 -fPIC -Wall -Wextra -Iinclude -c $< -o $@

$(OBJ)/%.cu.o: $(PKG)/csrc/cuda/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(ARCH) -std=c++17 -O3 -lineinfo -Xptxas -warn-spills -Xcompiler -fPIC -Iinclude -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS) -lm

oracle:
	$(MAKE) -C oracle

c-callers: $(LIB)
	gcc -std=c99 -O2 -Iinclude tests/c/reference_style_callers.c -o /tmp/vvdsp_callers -L$(PKG)/lib -lvvdsp_b200 -Wl,-rpath,$(abspath $(PKG)/lib) -lm
	gcc -std=gnu99 -O2 -Iinclude tests/c/perframe_latency.c -o /tmp/vvdsp_perframe_latency -L$(PKG)/lib -lvvdsp_b200 -Wl,-rpath,$(abspath $(PKG)/lib) -lm

clean:
	rm -rf $(OBJ) $(LIB)

.PHONY: all oracle c-callers clean
