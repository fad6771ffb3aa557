# Top-level build: the product library (sm_100a CUDA + C host shim) and the CPU checker.
#   make            -> obs-color-monitor_b200/lib/libscope_b200.so (+ libcm_shim.so), oracle/
#   make product    -> product libraries only
NVCC ?= /usr/local/cuda/bin/nvcc
PKG = obs-color-monitor_b200
LIBDIR = $(PKG)/lib
NVFLAGS = -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
          -Xcompiler -fPIC,-Wall -cudart static
# -Xptxas -v is noisy; `make PTXAS_V=1` shows registers / spills / smem per kernel
ifdef PTXAS_V
NVFLAGS += -Xptxas -v
endif

all: product oracle

product: $(LIBDIR)/libscope_b200.so $(LIBDIR)/libcm_shim.so

$(LIBDIR)/libscope_b200.so: Makefile $(PKG)/csrc/exports.map $(PKG)/csrc/scope_ffi.cu $(PKG)/csrc/scope_kernels.cuh $(PKG)/csrc/scope_kernels_experiments.cuh $(PKG)/csrc/scope_peer_reduce.cuh include/scope_ffi.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/scope_ffi.cu -Xlinker --version-script=$(PKG)/csrc/exports.map

$(LIBDIR)/libcm_shim.so: $(PKG)/csrc/cm_shim.c include/cm_shim.h include/scope_ffi.h $(LIBDIR)/libscope_b200.so
	gcc -std=gnu11 -O2 -g -fPIC -Wall -Wextra -shared -o $@ $(PKG)/csrc/cm_shim.c -Iinclude \
	    -L$(LIBDIR) -lscope_b200 -Wl,-rpath,'$$ORIGIN' -lpthread -lm

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(LIBDIR)
	$(MAKE) -C oracle clean
.PHONY: all product oracle clean

# A/B builds for kernel work (tools/run_ab.sh runs the GPU parity tests on, and benches, every
# variants_tmp/*.so through SCOPE_LIB).  `python tools/sass_budget.py variants_tmp/X.so` gives the static
# instruction budget of a variant's steady-state loop without a GPU.  Names ending in _x are built with
# SCOPE_EXPERIMENT: their two-plane (surface mode) rings do not fit, run_ab.sh skips those tests for them.
VARIANT = $(NVCC) $(NVFLAGS) -shared $(PKG)/csrc/scope_ffi.cu -Xlinker --version-script=$(PKG)/csrc/exports.map
VARIANTS = dephase wide_dephase wide_straight_dephase w8_dephase wide wide_straight immcoef w16n8_immcoef_r120_x w16n8_straight_immcoef_r120_x w8 w12n8_x w12n6_x w16n6_x w16n6_straight_x w16n8_r120_x w16n8_straight_r120_x w16n6_straight_r120_x r120 straight w8_straight w12n8_straight_x ballot w8_straight_ballot deepring nopipe rawflat w8_deepring nofaddr nodefer
FLAGS_w8 = -DSCOPE_TMA_WARPS=8
FLAGS_dephase = -DSCOPE_DEPHASE=1
FLAGS_wide_dephase = -DSCOPE_WIDE_FUSED=1 -DSCOPE_IMMCOEF=1 -DSCOPE_DEPHASE=1
FLAGS_wide_straight_dephase = -DSCOPE_WIDE_FUSED=1 -DSCOPE_IMMCOEF=1 -DSCOPE_STRAIGHT=1 -DSCOPE_DEPHASE=1
FLAGS_w8_dephase = -DSCOPE_TMA_WARPS=8 -DSCOPE_DEPHASE=1
FLAGS_immcoef = -DSCOPE_IMMCOEF=1
FLAGS_wide = -DSCOPE_WIDE_FUSED=1 -DSCOPE_IMMCOEF=1
FLAGS_wide_straight = -DSCOPE_WIDE_FUSED=1 -DSCOPE_IMMCOEF=1 -DSCOPE_STRAIGHT=1
FLAGS_w16n8_immcoef_r120_x = -DSCOPE_IMMCOEF=1 -DSCOPE_EXPERIMENT -DSCOPE_TILE_ROWS=128 -DSCOPE_MAXNREG=120
FLAGS_w16n8_straight_immcoef_r120_x = -DSCOPE_IMMCOEF=1 -DSCOPE_EXPERIMENT -DSCOPE_TILE_ROWS=128 -DSCOPE_STRAIGHT=1 -DSCOPE_MAXNREG=120
FLAGS_w12n8_x = -DSCOPE_EXPERIMENT -DSCOPE_TMA_WARPS=12 -DSCOPE_TILE_ROWS=96
FLAGS_w12n6_x = -DSCOPE_EXPERIMENT -DSCOPE_TMA_WARPS=12 -DSCOPE_TILE_ROWS=72
FLAGS_w16n6_x = -DSCOPE_EXPERIMENT -DSCOPE_TILE_ROWS=96
FLAGS_w16n6_straight_x = -DSCOPE_EXPERIMENT -DSCOPE_TILE_ROWS=96 -DSCOPE_STRAIGHT=1
FLAGS_w16n6_straight_r120_x = -DSCOPE_EXPERIMENT -DSCOPE_TILE_ROWS=96 -DSCOPE_STRAIGHT=1 -DSCOPE_MAXNREG=120
FLAGS_w16n8_r120_x = -DSCOPE_EXPERIMENT -DSCOPE_TILE_ROWS=128 -DSCOPE_MAXNREG=120
FLAGS_w16n8_straight_r120_x = -DSCOPE_EXPERIMENT -DSCOPE_TILE_ROWS=128 -DSCOPE_STRAIGHT=1 -DSCOPE_MAXNREG=120
FLAGS_r120 = -DSCOPE_MAXNREG=120
FLAGS_deepring = -DSCOPE_DEEP_RING=1
FLAGS_nopipe = -DSCOPE_PIPELINE=0
FLAGS_rawflat = -DSCOPE_RAWFLAT=1
FLAGS_straight = -DSCOPE_STRAIGHT=1
FLAGS_ballot = -DSCOPE_BALLOT=1
FLAGS_w8_straight_ballot = -DSCOPE_BALLOT=1 -DSCOPE_STRAIGHT=1 -DSCOPE_TMA_WARPS=8
FLAGS_w8_straight = -DSCOPE_STRAIGHT=1 -DSCOPE_TMA_WARPS=8
FLAGS_w12n8_straight_x = -DSCOPE_STRAIGHT=1 -DSCOPE_EXPERIMENT -DSCOPE_TMA_WARPS=12 -DSCOPE_TILE_ROWS=96
FLAGS_w8_deepring = -DSCOPE_TMA_WARPS=8 -DSCOPE_DEEP_RING=1
FLAGS_nofaddr = -DSCOPE_FADDR=0
FLAGS_nodefer = -DSCOPE_DEFER=0
variants: $(VARIANTS:%=variants_tmp/%.so)
variants_tmp/%.so: $(PKG)/csrc/scope_ffi.cu $(PKG)/csrc/scope_kernels.cuh $(PKG)/csrc/scope_kernels_experiments.cuh $(PKG)/csrc/scope_peer_reduce.cuh include/scope_ffi.h Makefile
	@mkdir -p variants_tmp
	$(VARIANT) $(FLAGS_$*) -o $@
.PHONY: variants
