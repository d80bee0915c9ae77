# Builds every native artefact in-tree (they travel to the GPU box with the snapshot; none is committed).
#   bcd_b200/libbcd_b200.so      C ABI (include/bcd_b200.h) + sm_100a kernels            [nvcc]
#   bcd_b200/libbcdcore_b200.so  bcd:: C++ API mirror (include/bcd/core/*.h) + synthetic-input generator [g++]
#   bcd_b200/libbcdsynth.so      the synthetic-input generator alone (no CUDA dependency)                [g++]
#   bin/bcd_cli, bin/raw2bcd     command-line front ends (EXR in, EXR out)                  [g++]
#   oracle/liboracle.so, oracle/_ref/libbcd_ref.so, oracle/_ref/libexr_ref.so   parity oracles (test infrastructure)   [gcc/g++]
NVCC ?= /usr/local/cuda/bin/nvcc
CXX = g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -lineinfo --fmad=false -std=c++17 -ccbin $(CXX) -Xcompiler -fPIC -Xptxas -v
CSRC = bcd_b200/csrc
KOBJ = $(CSRC)/k_distance.o $(CSRC)/k_misc.o $(CSRC)/k_marking.o $(CSRC)/k_filter.o $(CSRC)/k_eig_sub.o $(CSRC)/k_peer.o $(CSRC)/bcd_b200_capi.o $(CSRC)/bcd_b200_multi.o
HOSTSRC = $(wildcard bcd_b200/host/*.cpp) $(wildcard bcd_b200/io/*.cpp)
HOSTINC = -Iinclude -Iinclude/bcd/core -Iinclude/bcd/io -Ibcd_b200/io

all: bcd_b200/libbcd_b200.so bcd_b200/libbcdcore_b200.so bcd_b200/libbcdsynth.so bin/bcd_cli bin/raw2bcd oracle ref_frontends

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/kernels.h $(CSRC)/engine.h include/bcd_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

bcd_b200/libbcd_b200.so: $(KOBJ)
	$(NVCC) $(ARCH) -shared -ccbin $(CXX) -o $@ $(KOBJ) -lcudart

bcd_b200/libbcdcore_b200.so: $(HOSTSRC) $(wildcard include/bcd/core/*.h include/bcd/io/*.h bcd_b200/io/*.h) include/bcd_b200.h include/bcd_b200_host.h bcd_b200/libbcd_b200.so
	$(CXX) -std=c++11 -O2 -fopenmp -fPIC -shared $(HOSTINC) -o $@ $(HOSTSRC) \
		-Lbcd_b200 -lbcd_b200 -lz -Wl,-rpath,'$$ORIGIN'

# the synthetic-input generator G alone (bcd_synth_generate + SamplesAccumulator), with no dependency on the CUDA library:
# bench.py's reference arm generates its frame with it, so that nothing of the product path is mapped into that process
bcd_b200/libbcdsynth.so: bcd_b200/host/synth.cpp bcd_b200/host/SamplesAccumulator.cpp $(wildcard include/bcd/core/*.h) include/bcd_b200_host.h
	$(CXX) -std=c++11 -O2 -fopenmp -fPIC -shared -DBCD_SYNTH_ONLY $(HOSTINC) -o $@ bcd_b200/host/synth.cpp bcd_b200/host/SamplesAccumulator.cpp

# front ends (reference: src/cli/main.cpp, src/raw_converter/main.cpp)
bin/%: apps/%.cpp bcd_b200/libbcdcore_b200.so
	@mkdir -p bin
	$(CXX) -std=c++11 -O2 $(HOSTINC) -o $@ $< -Lbcd_b200 -lbcdcore_b200 -lbcd_b200 -Wl,-rpath,'$$ORIGIN/../bcd_b200'

oracle:
	$(MAKE) -C oracle all

# TEST INFRASTRUCTURE — the reference's own front-end sources (src/cli/main.cpp, src/raw_converter/main.cpp), UNMODIFIED,
# compiled against this repository's headers and linked against its libraries: the source-compatibility proof of the
# drop-in (tests/test_reference_frontends.py, tests/test_gpu_cli.py). Only where /root/reference exists; the binaries
# land in oracle/_ref/ (git-ignored, they travel to the GPU box). Eigen is on the include path because main.cpp includes
# <Eigen/Dense> without using it.
REF ?= /root/reference
ref_frontends: bcd_b200/libbcdcore_b200.so
	@if [ -d $(REF)/src/cli ]; then $(MAKE) --no-print-directory oracle/_ref/ref_bcd_cli oracle/_ref/ref_raw2bcd; \
	 else echo "ref_frontends: $(REF) not present; using prebuilt oracle/_ref/ref_bcd_cli if any"; fi

oracle/_ref/ref_bcd_cli: $(REF)/src/cli/main.cpp bcd_b200/libbcdcore_b200.so $(wildcard include/bcd/core/*.h include/bcd/io/*.h)
	@mkdir -p oracle/_ref
	$(CXX) -std=c++11 -O2 -w -include cmath $(HOSTINC) -I$(REF)/ext/eigen -o $@ $< -Lbcd_b200 -lbcdcore_b200 -lbcd_b200 -Wl,-rpath,'$$ORIGIN/../../bcd_b200'

oracle/_ref/ref_raw2bcd: $(REF)/src/raw_converter/main.cpp bcd_b200/libbcdcore_b200.so $(wildcard include/bcd/core/*.h include/bcd/io/*.h)
	@mkdir -p oracle/_ref
	$(CXX) -std=c++11 -O2 -w -include cmath $(HOSTINC) -o $@ $< -Lbcd_b200 -lbcdcore_b200 -lbcd_b200 -Wl,-rpath,'$$ORIGIN/../../bcd_b200'

# Guard of the bit-exact pair-term kernel (k_distance.cu): the reference rounds ny*b1 and nx*b2 separately and divides with
# IEEE division. ptxas must therefore (a) keep the packed products as FMUL2 / FFMA2-with-zero-addend instead of contracting
# them into the following subtraction, (b) keep div.rn's refinement (5 packed FMA-pipe steps behind the two MUFU.RCP of the
# negated denominators) and (c) use the FSET masks of the contribution test. Counts below are those of the walk loops.
check-sass: $(CSRC)/k_distance.o
	@cuobjdump -sass -fun '_ZN4bcdk19k_pair_terms_sparseILi32ELi4ELi4ELi2EEEvPKfS2_NS_9LevelGeomEPfPh' $< > /tmp/bcd_b200_k1a.sass
	@n_rcpneg=$$(grep -c "MUFU.RCP R[0-9]*, -R" /tmp/bcd_b200_k1a.sass); n_fset=$$(grep -c "FSET.BF.GTU" /tmp/bcd_b200_k1a.sass); \
	 n_fmul2=$$(grep -c "FMUL2" /tmp/bcd_b200_k1a.sass); n_ffma2=$$(grep -c "FFMA2" /tmp/bcd_b200_k1a.sass); \
	 echo "k_pair_terms_sparse<32,4,4,2>: MUFU.RCP(-x) $$n_rcpneg, FSET.BF.GTU $$n_fset, FMUL2 $$n_fmul2, FFMA2 $$n_ffma2"; \
	 test $$n_rcpneg -ge 8 && test $$n_fset -ge 8 && test $$n_fmul2 -ge 8 && test $$n_ffma2 -ge 20 || (echo "check-sass: unexpected code shape in the pair-term walk" && exit 1)

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.log bcd_b200/*.so bin/bcd_cli bin/raw2bcd
	$(MAKE) -C oracle clean

.PHONY: all oracle ref_frontends clean check-sass
