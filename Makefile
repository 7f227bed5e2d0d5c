# Builds libcimq.so (the C-ABI CUDA library) in-tree for sm_100a.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall -Xptxas -v $(if $(TIMERS),-DCIMQ_TIMERS=1) $(EXTRA)
SRC_DIR   := cim_quantization_b200/csrc
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(SRCS:.cu=.o)
HDRS      := $(wildcard $(SRC_DIR)/*.cuh) include/cimq.h
LIB       := cim_quantization_b200/libcimq.so

all: $(LIB)

$(SRC_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(@:.o=.ptxas.log) || (cat $(@:.o=.ptxas.log); exit 1)

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(LIB) $(SRC_DIR)/*.ptxas.log

.PHONY: all clean
