# Builds libwebradio_b200.so (hand-written sm_100a CUDA behind the C ABI of include/webradio_b200.h)
# and the test harness that drives the DspBlock drop-in classes.  nvcc cross-compiles without a GPU.
#
#   make            -> webradio_b200/libwebradio_b200.so
#   make harness    -> tests/harness/libwr_blocks_harness.so  (C++ drop-in blocks + graph driver)
#   make oracle     -> oracle/libwr_oracle.so and, where /root/reference is mounted, oracle/_ref/
NVCC     ?= nvcc
CXX      ?= g++
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude -Iwebradio_b200/csrc
# sample path: every product/sum is an explicit _rn intrinsic; -fmad=false is belt and braces
PARITY   := -fmad=false
CSRC     := webradio_b200/csrc
OBJ      := build/wr_bank.o build/wr_stage.o build/wr_spectrum.o build/wr_upload.o build/wr_host.o
LIB      := webradio_b200/libwebradio_b200.so
HARNESS  := tests/harness/libwr_blocks_harness.so
BLOCKSRC := webradio_b200/dsp/dspblock.cxx webradio_b200/dsp/downconverter.cxx webradio_b200/dsp/lowpass.cxx \
            webradio_b200/dsp/demodulator.cxx webradio_b200/io/spectrumsink.cxx webradio_b200/dsp/gpubank.cxx

.PHONY: all lib lib-exp harness harness-mock dropin dropin-mock asan-check tsan-check oracle check tools clean
all: lib
lib: $(LIB)

build/wr_bank.o: $(CSRC)/wr_bank.cu $(wildcard $(CSRC)/*.cuh) $(CSRC)/wr_common.h include/webradio_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(PARITY) -c $< -o $@
build/wr_stage.o: $(CSRC)/wr_stage.cu $(wildcard $(CSRC)/*.cuh) $(CSRC)/wr_common.h include/webradio_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(PARITY) -c $< -o $@
build/wr_spectrum.o: $(CSRC)/wr_spectrum.cu $(wildcard $(CSRC)/*.cuh) $(CSRC)/wr_common.h include/webradio_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@
build/wr_upload.o: $(CSRC)/wr_upload.cu $(wildcard $(CSRC)/*.cuh) $(CSRC)/wr_common.h include/webradio_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@
build/wr_host.o: $(CSRC)/wr_host.cpp $(CSRC)/wr_common.h include/webradio_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xcompiler -ffp-contract=off -x cu -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart_static -lpthread -ldl -lrt

# Experimental twin of the library for A/B runs on a GPU box (never loaded unless WEBRADIO_B200_LIB points at it):
# same sources, experiment macros on.  EXPFLAGS picks the experiment(s).
EXPFLAGS ?= -DWR_EXP_FIR_SLEEP=200
LIB_EXP  := webradio_b200/libwebradio_b200_exp.so
lib-exp:
	@mkdir -p build/exp
	$(NVCC) $(NVFLAGS) $(PARITY) $(EXPFLAGS) -c $(CSRC)/wr_bank.cu -o build/exp/wr_bank.o
	$(NVCC) $(ARCH) -shared -o $(LIB_EXP) build/exp/wr_bank.o build/wr_stage.o build/wr_spectrum.o build/wr_upload.o build/wr_host.o -lcudart_static -lpthread -ldl -lrt

harness: $(HARNESS)
$(HARNESS): tests/harness/graph_harness.cxx $(BLOCKSRC) $(wildcard webradio_b200/dsp/*.h webradio_b200/io/*.h) $(LIB)
	$(CXX) -std=c++11 -O2 -fPIC -Wall -shared -Iinclude -Iwebradio_b200 -Iwebradio_b200/dsp -Iwebradio_b200/io \
	  -o $@ tests/harness/graph_harness.cxx $(BLOCKSRC) -Lwebradio_b200 -lwebradio_b200 -Wl,-Bsymbolic -Wl,-rpath,'$$ORIGIN/../../webradio_b200' -lpthread

# TEST INFRASTRUCTURE: the same drop-in blocks and graph driver over a CPU stand-in for the device
# entry points (tests/harness/mock_capi.cxx, arithmetic by the oracle) -- exercises the blocks'
# HOST logic in the CPU test suite.  The product's own host cold path (wr_host.o) is linked in.
HARNESS_MOCK := tests/harness/libwr_blocks_harness_mock.so
CUDA_LIB ?= /usr/local/cuda/lib64
harness-mock: $(HARNESS_MOCK)
$(HARNESS_MOCK): tests/harness/graph_harness.cxx tests/harness/mock_capi.cxx $(BLOCKSRC) $(wildcard webradio_b200/dsp/*.h webradio_b200/io/*.h) build/wr_host.o oracle/wr_oracle.c oracle/wr_oracle.h
	$(MAKE) -s -C oracle port
	$(CXX) -std=c++11 -O2 -fPIC -Wall -shared -Iinclude -Iwebradio_b200 -Iwebradio_b200/dsp -Iwebradio_b200/io \
	  -o $@ tests/harness/graph_harness.cxx tests/harness/mock_capi.cxx $(BLOCKSRC) build/wr_host.o \
	  -Loracle -lwr_oracle -L$(CUDA_LIB) -lcudart_static -Wl,-Bsymbolic -Wl,-rpath,'$$ORIGIN/../../oracle' -lpthread -ldl -lrt

# Host logic of the drop-in blocks under AddressSanitizer + UBSan (stand-in back-end, no GPU)
# the distribution's compiler: it ships the sanitizer runtimes (a toolchain under /opt may not)
ASAN_CC  ?= $(shell command -v /usr/bin/gcc || command -v gcc)
ASAN_CXX ?= $(shell command -v /usr/bin/g++ || command -v g++)
asan-check: build/wr_host.o
	@mkdir -p build
	$(ASAN_CC) -std=c11 -O1 -g -ffp-contract=off -fsanitize=address,undefined -c oracle/wr_oracle.c -o build/wr_oracle_asan.o
	$(ASAN_CXX) -std=c++11 -O1 -g -DWR_QUIET_DEBUG -fsanitize=address,undefined -fno-omit-frame-pointer -Iinclude -Iwebradio_b200 \
	  -Iwebradio_b200/dsp -Iwebradio_b200/io -o build/host_scenario_asan tests/harness/host_scenario.cxx \
	  tests/harness/graph_harness.cxx tests/harness/mock_capi.cxx $(BLOCKSRC) build/wr_oracle_asan.o build/wr_host.o \
	  -L$(CUDA_LIB) -lcudart_static -lpthread -ldl -lrt -lm
	ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=halt_on_error=1 ./build/host_scenario_asan
	@if [ -f "$(REF)/src/radio.cxx" ]; then \
	  $(ASAN_CXX) -std=c++11 -O1 -g -DWR_QUIET_DEBUG -fsanitize=address,undefined -fno-omit-frame-pointer -Iinclude \
	    -Itests/harness/stubs -Iwebradio_b200 -Iwebradio_b200/dsp -Iwebradio_b200/io -I$(REF)/src -I$(REF)/src/io \
	    -o build/radio_scenario_asan tests/harness/radio_scenario.cxx $(REF)/src/radio.cxx tests/harness/radio_dropin.cxx \
	    tests/harness/mock_capi.cxx $(BLOCKSRC) build/wr_oracle_asan.o build/wr_host.o \
	    -L$(CUDA_LIB) -lcudart_static -lpthread -ldl -lrt -lm && \
	  ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=halt_on_error=1 ./build/radio_scenario_asan; \
	else echo "reference tree not mounted: radio glue scenario skipped"; fi

# ... and under ThreadSanitizer: the DSP thread against two threads doing what the REST handlers do
tsan-check: build/wr_host.o
	@mkdir -p build
	$(ASAN_CC) -std=c11 -O1 -g -ffp-contract=off -fsanitize=thread -c oracle/wr_oracle.c -o build/wr_oracle_tsan.o
	$(ASAN_CXX) -std=c++11 -O1 -g -DWR_QUIET_DEBUG -fsanitize=thread -fno-omit-frame-pointer -Iinclude -Iwebradio_b200 \
	  -Iwebradio_b200/dsp -Iwebradio_b200/io -o build/host_threads_tsan tests/harness/host_threads.cxx \
	  tests/harness/graph_harness.cxx tests/harness/mock_capi.cxx $(BLOCKSRC) build/wr_oracle_tsan.o build/wr_host.o \
	  -L$(CUDA_LIB) -lcudart_static -lpthread -ldl -lrt -lm
	TSAN_OPTIONS=halt_on_error=1 ./build/host_threads_tsan

# The reference's own graph glue (src/radio.cxx, UNMODIFIED, compiled where it lies) linked against
# the drop-in blocks.  Only buildable where the reference tree is mounted; the .so travels.
REF ?= /root/reference
DROPIN := tests/harness/libwr_radio_dropin.so
dropin: $(LIB)
	@if [ -f "$(REF)/src/radio.cxx" ]; then \
	  $(CXX) -std=c++11 -O2 -fPIC -Wall -shared -DWR_QUIET_DEBUG -Iinclude -Itests/harness/stubs -Iwebradio_b200 \
	    -Iwebradio_b200/dsp -Iwebradio_b200/io -I$(REF)/src -I$(REF)/src/io \
	    -o $(DROPIN) $(REF)/src/radio.cxx tests/harness/radio_dropin.cxx $(BLOCKSRC) \
	    -Lwebradio_b200 -lwebradio_b200 -Wl,-Bsymbolic -Wl,-rpath,'$$ORIGIN/../../webradio_b200' -lpthread && echo "built $(DROPIN)"; \
	else echo "reference tree not mounted: keeping prebuilt $(DROPIN) (if any)"; fi

# ... and the same glue over the CPU stand-in of tests/harness/mock_capi.cxx (CPU test suite)
DROPIN_MOCK := tests/harness/libwr_radio_dropin_mock.so
dropin-mock: build/wr_host.o
	@if [ -f "$(REF)/src/radio.cxx" ]; then \
	  $(MAKE) -s -C oracle port && \
	  $(CXX) -std=c++11 -O2 -fPIC -Wall -shared -DWR_QUIET_DEBUG -Iinclude -Itests/harness/stubs -Iwebradio_b200 \
	    -Iwebradio_b200/dsp -Iwebradio_b200/io -I$(REF)/src -I$(REF)/src/io \
	    -o $(DROPIN_MOCK) $(REF)/src/radio.cxx tests/harness/radio_dropin.cxx tests/harness/mock_capi.cxx $(BLOCKSRC) build/wr_host.o \
	    -Loracle -lwr_oracle -L$(CUDA_LIB) -lcudart_static -Wl,-Bsymbolic -Wl,-rpath,'$$ORIGIN/../../oracle' -lpthread -ldl -lrt && echo "built $(DROPIN_MOCK)"; \
	else echo "reference tree not mounted: keeping prebuilt $(DROPIN_MOCK) (if any)"; fi

oracle:
	$(MAKE) -C oracle port ref

# everything that can be checked without a GPU
check: lib harness harness-mock oracle
	python -m pytest tests -q -m "not gpu"

# stand-alone micro-benchmarks (run on the GPU box; build/ travels with gpurun)
tools: build/ubench_copy build/ubench_dsmem
build/ubench_copy: tools/ubench_copy.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O2 -o $@ $<
build/ubench_dsmem: tools/ubench_dsmem.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -o $@ $<

clean:
	rm -rf build $(LIB) $(HARNESS) $(HARNESS_MOCK)
