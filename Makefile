# Builds every native artefact in-tree (they travel to the GPU box with the snapshot; none is committed).
#   bcd_b200/libbcd_b200.so      C ABI (include/bcd_b200.h) + sm_100a kernels            [nvcc]
#   bcd_b200/libbcdcore_b200.so  bcd:: C++ API mirror (include/bcd/core/*.h) + synthetic-input generator [g++]
#   oracle/liboracle.so, oracle/_ref/libbcd_ref.so   parity oracles (test infrastructure)   [gcc/g++]
NVCC ?= /usr/local/cuda/bin/nvcc
CXX = g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -lineinfo --fmad=false -std=c++17 -ccbin $(CXX) -Xcompiler -fPIC -Xptxas -v
CSRC = bcd_b200/csrc
KOBJ = $(CSRC)/k_distance.o $(CSRC)/k_misc.o $(CSRC)/k_marking.o $(CSRC)/k_filter.o $(CSRC)/k_eig_sub.o $(CSRC)/bcd_b200_capi.o
HOSTSRC = $(wildcard bcd_b200/host/*.cpp)

all: bcd_b200/libbcd_b200.so bcd_b200/libbcdcore_b200.so oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/kernels.h include/bcd_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

bcd_b200/libbcd_b200.so: $(KOBJ)
	$(NVCC) $(ARCH) -shared -ccbin $(CXX) -o $@ $(KOBJ) -lcudart

bcd_b200/libbcdcore_b200.so: $(HOSTSRC) $(wildcard include/bcd/core/*.h) include/bcd_b200.h bcd_b200/libbcd_b200.so
	$(CXX) -std=c++11 -O2 -fopenmp -fPIC -shared -Iinclude -Iinclude/bcd/core -o $@ $(HOSTSRC) \
		-Lbcd_b200 -lbcd_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.log bcd_b200/*.so
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
