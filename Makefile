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

$(LIBDIR)/libscope_b200.so: Makefile $(PKG)/csrc/exports.map $(PKG)/csrc/scope_ffi.cu $(PKG)/csrc/scope_kernels.cuh include/scope_ffi.h
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

# A/B builds for kernel work (tools/run_ab.sh benches every variants_tmp/*.so through SCOPE_LIB)
VARIANT = $(NVCC) $(NVFLAGS) -shared $(PKG)/csrc/scope_ffi.cu -Xlinker --version-script=$(PKG)/csrc/exports.map
variants:
	@mkdir -p variants_tmp
	$(VARIANT) -DSCOPE_LDSM=0 -DSCOPE_XORSWZ=0 -DSCOPE_DEFER=0 -DSCOPE_FADDR=0 -DSCOPE_FAST_EMIT=0 -o variants_tmp/base.so
	$(VARIANT) -DSCOPE_FADDR=0 -o variants_tmp/nofaddr.so
	$(VARIANT) -DSCOPE_FAST_EMIT=0 -o variants_tmp/noemit.so
	$(VARIANT) -DSCOPE_MAX_CHUNK=20 -o variants_tmp/chunk20.so
	$(VARIANT) -DSCOPE_DEFER=0 -o variants_tmp/nodefer.so
	$(VARIANT) -DSCOPE_RAWFLAT=1 -o variants_tmp/rawflat.so
.PHONY: variants
