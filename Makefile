# Builds libdamsm_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree.
# Variant builds for A/B timing: make OUT=tools/_ab/libX.so OBJDIR=build/X EXTRA=-DSOMETHING
NVCC ?= /usr/local/cuda/bin/nvcc
PKG := t2i_clip-gan_b200
SRC := $(wildcard $(PKG)/csrc/*.cu)
HDR := $(wildcard $(PKG)/csrc/*.cuh) include/damsm_b200.h
OUT ?= $(PKG)/libdamsm_b200.so
OBJDIR ?= build/obj
OBJ := $(patsubst $(PKG)/csrc/%.cu,$(OBJDIR)/%.o,$(SRC))
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
           -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $(EXTRA)

all: $(OUT)

$(OBJDIR)/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(OUT): $(OBJ)
	@mkdir -p $(dir $(OUT))
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJ)

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o /tmp/damsm_ptxas.so $(SRC) 2>&1 | grep -E "Compiling|registers|spill"

clean:
	rm -rf $(OUT) $(OBJDIR)
