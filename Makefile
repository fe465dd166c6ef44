# Builds libtokb200.so (sm_100a CUDA kernels + C ABI), the GPU self-test binary and the C oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v
CSRC      := torchok_b200/csrc
SRCS      := $(CSRC)/tok_conv.cu $(CSRC)/tok_conv2.cu $(CSRC)/tok_conv3.cu $(CSRC)/tok_api.cu $(CSRC)/tok_elem.cu $(CSRC)/tok_bn2.cu $(CSRC)/tok_retrieval.cu $(CSRC)/tok_retrieval2.cu $(CSRC)/tok_heads.cu $(CSRC)/tok_seg.cu $(CSRC)/tok_swin.cu $(CSRC)/tok_comm.cu $(CSRC)/tok_ocr.cu
OBJS      := $(SRCS:.cu=.o)
LIB       := torchok_b200/libtokb200.so

all: $(LIB) tests/gpu/tok_selftest

%.o: %.cu $(CSRC)/tok_bnfin.cuh $(CSRC)/tok_topk.cuh $(CSRC)/tok_optim.cuh $(CSRC)/tok_ptx.cuh $(CSRC)/tok_pair.cuh $(CSRC)/tok_conv.cuh $(CSRC)/tok_internal.h include/tokb200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static

tests/gpu/tok_selftest: tests/gpu/tok_selftest.cu $(LIB) include/tokb200.h
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ $< -Ltorchok_b200 -ltokb200 -Xlinker -rpath -Xlinker '$$ORIGIN/../../torchok_b200' -cudart static

# bring-up probe of the cta_group::2 UMMA path (stand-alone; not part of `all`)
probe: tests/gpu/gemm2cta_probe
tests/gpu/gemm2cta_probe: tests/gpu/gemm2cta_probe.cu $(CSRC)/tok_ptx.cuh $(CSRC)/tok_pair.cuh
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -I $(CSRC) -o $@ $< -lcuda

clean:
	rm -f $(OBJS) $(CSRC)/*.log $(LIB) tests/gpu/tok_selftest tests/gpu/gemm2cta_probe

.PHONY: all clean probe
