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

$(LIBDIR)/libscope_b200.so: Makefile $(PKG)/csrc/exports.map $(PKG)/csrc/scope_ffi.cu $(PKG)/csrc/scope_kernels.cuh $(PKG)/csrc/scope_kernels_experiments.cuh $(PKG)/csrc/scope_fused_v3.cuh $(PKG)/csrc/scope_peer_reduce.cuh include/scope_ffi.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/scope_ffi.cu -Xlinker --version-script=$(PKG)/csrc/exports.map

$(LIBDIR)/libcm_shim.so: $(PKG)/csrc/cm_shim.c include/cm_shim.h include/scope_ffi.h $(LIBDIR)/libscope_b200.so
	gcc -std=gnu11 -O2 -g -fPIC -Wall -Wextra -shared -o $@ $(PKG)/csrc/cm_shim.c -Iinclude -I/usr/local/cuda/include \
	    -L$(LIBDIR) -lscope_b200 -Wl,-rpath,'$$ORIGIN' -lpthread -lm -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(LIBDIR)
	$(MAKE) -C oracle clean
.PHONY: all product oracle clean

# A/B builds for kernel work: `SCOPE_LIB=$PWD/variants_tmp/X.so python bench.py ...` (tools/gpu_call30 ... 41.sh bench every
# variants_tmp/*.so they find).  What round 2 measured with them: profiles/r02/ab_v3/ and profiles/r02/c30 ... c41.
#   v3off            the general kernel serves the headline combination too (the round-1 path)
#   v3_base          scope_fused_kernel_v3 as it stood at 42.9 %: per-visit checks, deferred look at the vectorscope adds, LDC constants
#   v3_lean_defer / v3_lean_ldc / v3_lean_w24   lean visits with the deferred look / with LDC constants / 24 warps at 80 registers
#   v3_lean_narrow   one level per end-of-strip step instead of four
#   v3_noaffine      one work counter over the batch instead of one per frame
#   v3_nobg          without the background-lane rule for screen content
#   v3_pf0 / _pf3 / _pf12 / _pf20   L2 prefetch distance in tiles (shipped: 6)
#   v3_w23s4, v3_w25s3, v3_w26s3   consumer warps x ring stages (shipped: 27 x 3)
#   v3_scalar        scalar FFMA / FADD instead of the f32x2 forms
#   v3_nop           DIAGNOSTIC (results wrong by construction): ring, end-of-strip write-out and flushes only, no accumulation
#   v3_noload        DIAGNOSTIC: no TMA at all, the consumers accumulate whatever the stages hold
#   v3_d1/_d2/_d4/_d6  DIAGNOSTIC: no end-of-strip write-out / not even its barriers / no vectorscope flush / neither
#   v3_s1/_s2/_s3    DIAGNOSTIC: without the column-bin adds / the vectorscope adds / both (addresses still computed)
#   v3_red           DIAGNOSTIC: vectorscope adds without return value (saturation of flat content then wrong)
#   w8, straight, nopipe: build flags of the general kernel that tests/test_kernel_emulation.py still covers
VARIANT = $(NVCC) $(NVFLAGS) -shared $(PKG)/csrc/scope_ffi.cu -Xlinker --version-script=$(PKG)/csrc/exports.map
VARIANTS = v3_base v3_lean_defer v3_lean_ldc v3_lean_w24 v3_lean_narrow v3_noaffine v3_nobg v3_pf12 v3_pf20 v3_pf3 v3_d1 v3_d2 v3_d4 v3_d6 v3_w23s4 v3_w25s3 v3_w26s3 v3_w27pf0 v3off v3_scalar v3_pf0 v3_nop v3_noload v3_s1 v3_s2 v3_s3 v3_red w8 straight nopipe
FLAGS_v3off = -DSCOPE_V3=0
FLAGS_v3_base = -DSCOPE_V3_LEAN=0 -DSCOPE_V3_RESOLVE_NOW=0 -DSCOPE_V3_SMEM_CONSTS=0 -DSCOPE_V3_WIDE_EMIT=0 -DSCOPE_V3_FRAME_AFFINE=0 -DSCOPE_V3_BG_SKIP=0
FLAGS_v3_lean_defer = -DSCOPE_V3_RESOLVE_NOW=0
FLAGS_v3_lean_w24 = -DSCOPE_V3_WARPS=24
FLAGS_v3_lean_ldc = -DSCOPE_V3_SMEM_CONSTS=0
FLAGS_v3_lean_narrow = -DSCOPE_V3_WIDE_EMIT=0
FLAGS_v3_noaffine = -DSCOPE_V3_FRAME_AFFINE=0
FLAGS_v3_nobg = -DSCOPE_V3_BG_SKIP=0
FLAGS_v3_pf3 = -DSCOPE_V3_L2_AHEAD=3
FLAGS_v3_d1 = -DSCOPE_V3_DIAG=1
FLAGS_v3_d2 = -DSCOPE_V3_DIAG=2
FLAGS_v3_d4 = -DSCOPE_V3_DIAG=4
FLAGS_v3_d6 = -DSCOPE_V3_DIAG=6
FLAGS_v3_pf12 = -DSCOPE_V3_L2_AHEAD=12
FLAGS_v3_pf20 = -DSCOPE_V3_L2_AHEAD=20
FLAGS_v3_w23s4 = -DSCOPE_V3_WARPS=23 -DSCOPE_V3_STAGES=4
FLAGS_v3_w25s3 = -DSCOPE_V3_WARPS=25 -DSCOPE_V3_STAGES=3
FLAGS_v3_w26s3 = -DSCOPE_V3_WARPS=26 -DSCOPE_V3_STAGES=3
FLAGS_v3_w27pf0 = -DSCOPE_V3_L2_AHEAD=0
FLAGS_v3_scalar = -DSCOPE_V3_FFMA2=0
FLAGS_v3_pf0 = -DSCOPE_V3_L2_AHEAD=0
FLAGS_v3_nop = -DSCOPE_V3_NOP
FLAGS_v3_noload = -DSCOPE_V3_NOLOAD
FLAGS_v3_s1 = -DSCOPE_V3_SKIP=1
FLAGS_v3_s2 = -DSCOPE_V3_SKIP=2
FLAGS_v3_s3 = -DSCOPE_V3_SKIP=3
FLAGS_v3_red = -DSCOPE_V3_SKIP=4
FLAGS_w8 = -DSCOPE_TMA_WARPS=8
FLAGS_straight = -DSCOPE_STRAIGHT=1
FLAGS_nopipe = -DSCOPE_PIPELINE=0
variants: $(VARIANTS:%=variants_tmp/%.so)
variants_tmp/%.so: $(PKG)/csrc/scope_ffi.cu $(PKG)/csrc/scope_kernels.cuh $(PKG)/csrc/scope_fused_v3.cuh $(PKG)/csrc/scope_kernels_experiments.cuh $(PKG)/csrc/scope_peer_reduce.cuh include/scope_ffi.h Makefile
	@mkdir -p variants_tmp
	$(VARIANT) $(FLAGS_$*) -o $@
.PHONY: variants
