# Builds libdiagon_b200.so (CUDA engine + C++20 host layer + C ABI) for sm_100a, in-tree.
#   make            -> diagon_b200/libdiagon_b200.so
#   make oracle     -> oracle/_build/liboracle.so (+ oracle/_ref when /root/reference is present)
NVCC     := nvcc -ccbin /usr/bin/g++
CXX      := /usr/bin/g++
CUDA_HOME ?= /usr/local/cuda
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v -maxrregcount=128 $(EXTRA_NVFLAGS)
# host arithmetic must match the reference's IEEE build: no fast-math, no FMA contraction
CXXFLAGS := -std=c++20 -O2 -fPIC -Wall -Wextra -ffp-contract=off -fno-fast-math -pthread
OUT      := diagon_b200/libdiagon_b200.so
BUILD    := build

HOST_SRCS := diagon_b200/host/host_index.cpp diagon_b200/host/segment_reader.cpp diagon_b200/host/search.cpp diagon_b200/host/c_api.cpp
HOST_OBJS := $(patsubst diagon_b200/host/%.cpp,$(BUILD)/%.o,$(HOST_SRCS))
HOST_HDRS := $(wildcard diagon_b200/host/*.h) $(wildcard include/*.h)

all: $(OUT)

$(BUILD)/%.o: diagon_b200/host/%.cpp $(HOST_HDRS)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

# the flags are a dependency too: `make EXTRA_NVFLAGS=-DDGPU_CHECK` followed by a plain `make` must rebuild
$(BUILD)/.nvflags: FORCE
	@mkdir -p $(BUILD)
	@echo '$(NVFLAGS)' | cmp -s - $@ || echo '$(NVFLAGS)' > $@

FORCE:

$(BUILD)/engine.o: diagon_b200/csrc/engine.cu $(wildcard diagon_b200/csrc/*.cuh) include/dgpu_engine.h Makefile $(BUILD)/.nvflags
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/ptxas.log || (cat $(BUILD)/ptxas.log; false)

$(OUT): $(HOST_OBJS) $(BUILD)/engine.o
	$(CXX) -shared -o $@ $^ -L$(CUDA_HOME)/lib64 -lcudart_static -ldl -lrt -lpthread

oracle:
	$(MAKE) -C oracle port
	$(MAKE) -C oracle -j8 ref

clean:
	rm -rf $(BUILD) $(OUT)

.PHONY: all oracle clean FORCE
