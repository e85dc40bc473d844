# Builds libfsb.so (the C-ABI library, sm_100a only) in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
PKG := fish_speech_rs_b200
CSRC := $(PKG)/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function --expt-relaxed-constexpr
SRCS := $(CSRC)/fsb_common.cu $(CSRC)/fsb_lm.cu $(CSRC)/fsb_lm_mega_bf16.cu $(CSRC)/fsb_lm_mega_f32.cu $(CSRC)/fsb_lm_mega1_bf16.cu $(CSRC)/fsb_lm_mega1_f32.cu $(CSRC)/fsb_lm_megab.cu $(CSRC)/fsb_tc_gemm.cu $(CSRC)/fsb_tc_conv.cu $(CSRC)/fsb_codec.cu
OBJS := $(SRCS:.cu=.o)
HDRS := $(wildcard $(CSRC)/*.cuh) include/fsb.h

all: $(PKG)/libfsb.so

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(PKG)/libfsb.so: $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(PKG)/libfsb.so
.PHONY: all clean
