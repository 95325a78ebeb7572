# Builds the C-ABI shared library (hand-written sm_100a CUDA) in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -DSSR_WARPLOCAL -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas --expt-relaxed-constexpr
SRC := ssr_eval_b200/csrc
OBJ := build/obj
LIB := ssr_eval_b200/lib/libssr_b200.so
SRCS := $(SRC)/stft_metrics.cu $(SRC)/resample.cu $(SRC)/stft_lowpass.cu $(SRC)/stft_splice.cu $(SRC)/sosfiltfilt.cu $(SRC)/pcm.cu $(SRC)/stft_lowpass_dense.cu $(SRC)/xcorr_align.cu
OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(SRCS))
HDRS := $(wildcard $(SRC)/*.cuh) $(SRC)/stft_tables.hpp $(SRC)/resample_tables.hpp include/ssr_b200.h Makefile

all: $(LIB)

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(PTXAS_V) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -cudart static

# host-side emulation harness for the FFT core (no GPU needed)
build/host_emul: tests/host_emul.cu $(SRC)/fft_core.cuh $(SRC)/stft_tables.hpp $(SRC)/k1_map.cuh $(SRC)/resample_tables.hpp
	@mkdir -p build
	$(NVCC) -O2 -std=c++17 --expt-relaxed-constexpr -o $@ $<

# A/B variants (development): make variant VARIANT=name VARIANT_FLAGS="-DSSR_..." -> build/variants/libssr_b200_name.so
# (timed against the in-tree build by tools/ab_lib.py in one process)
VARIANT ?= x
variant:
	@mkdir -p build/variants/obj_$(VARIANT)
	@for f in $(SRCS); do b=$$(basename $$f .cu); \
	  $(NVCC) $(NVFLAGS) $(VARIANT_FLAGS) -c $$f -o build/variants/obj_$(VARIANT)/$$b.o & done; wait
	$(NVCC) -shared $(ARCH) -o build/variants/libssr_b200_$(VARIANT).so build/variants/obj_$(VARIANT)/*.o -cudart static

clean:
	rm -rf build $(LIB)

.PHONY: all clean variant
