#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. -DFLAG2=.."  -> profiles/scratch/NAME.so  (only ctc_loss_fast.cu is recompiled)
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/scratch
B=silent-speech-decoder_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $2 \
  -I include -I silent-speech-decoder_b200/csrc -c silent-speech-decoder_b200/csrc/ctc_loss_fast.cu -o profiles/scratch/$1.o 2>&1 | grep -i "error" || true
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o profiles/scratch/$1.so $B/greedy.o $B/beam.o $B/ctc_loss.o $B/edit.o $B/aux.o $B/capi.o profiles/scratch/$1.o
