#!/bin/bash
# build_beam_variant.sh NAME "-DFLAG=.."  -> profiles/scratch/NAME.so  (only beam.cu is recompiled)
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/scratch
B=silent-speech-decoder_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $2 \
  -I include -I silent-speech-decoder_b200/csrc -c silent-speech-decoder_b200/csrc/beam.cu -o profiles/scratch/$1.o 2>&1 | grep -E "error|beam_kernelILb1|Used|spill" | grep -A2 ILb1 | tail -3 || true
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o profiles/scratch/$1.so $B/greedy.o profiles/scratch/$1.o $B/ctc_loss.o $B/ctc_loss_fast.o $B/edit.o $B/aux.o $B/capi.o
