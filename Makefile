# Top-level build: the product library (sm_100a only), the test oracle, and the drop-in client.
#   make            -> ddc_svd_b200/libsvdgpu.so
#   make oracle     -> oracle/libddcoracle.so (+ oracle/_ref/libddcref.so when /root/reference exists)
#   make dropin     -> build/test-whole-svd from the reference's UNMODIFIED driver source, and build/dropin_check
#                      (tests/dropin_check.c: the same flow with the driver's dormant check enabled)
NVCC      ?= nvcc
CC         = gcc
ARCH       = -gencode arch=compute_100a,code=sm_100a
NVFLAGS    = $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
CFLAGS     = -std=gnu99 -O2 -fPIC -Wall
PKG        = ddc_svd_b200
CU_SRCS    = $(wildcard $(PKG)/csrc/*.cu)
C_SRCS     = $(wildcard $(PKG)/host/*.c)
CU_OBJS    = $(patsubst $(PKG)/csrc/%.cu,build/obj/%.o,$(CU_SRCS))
C_OBJS     = $(patsubst $(PKG)/host/%.c,build/obj/host_%.o,$(C_SRCS))
HDRS       = $(wildcard $(PKG)/csrc/*.cuh) $(wildcard include/*.h)
LIB        = $(PKG)/libsvdgpu.so
REF       ?= /root/reference

all: $(LIB)

build/obj/%.o: $(PKG)/csrc/%.cu $(HDRS)
	@mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/obj/host_%.o: $(PKG)/host/%.c $(HDRS)
	@mkdir -p build/obj
	$(CC) $(CFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS) $(C_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart -ldl -lpthread

oracle:
	$(MAKE) -C oracle all

# The reference's own driver, compiled as is against our headers and library.
# It is fed through stdin so that its #include "cl-helper.h" / "svd_gpu.h" / "matrix_helper.h"
# resolve to include/ here instead of the reference's own directory.
dropin: $(LIB)
	@mkdir -p build
	@if [ -f $(REF)/test-whole-svd.c ]; then \
	  $(CC) -std=gnu99 -O2 -Iinclude -o build/test-whole-svd -x c - < $(REF)/test-whole-svd.c \
	        -L$(PKG) -lsvdgpu -Wl,-rpath,'$$ORIGIN/../$(PKG)' -lm && echo built build/test-whole-svd; \
	else echo "dropin: $(REF)/test-whole-svd.c not present"; fi
	$(CC) -std=gnu99 -O2 -Iinclude -o build/dropin_check tests/dropin_check.c \
	      -L$(PKG) -lsvdgpu -Wl,-rpath,'$$ORIGIN/../$(PKG)' -lm && echo built build/dropin_check

clean:
	rm -rf build $(LIB)
.PHONY: all oracle dropin clean
