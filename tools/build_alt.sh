#!/bin/bash
# Build an alternative libctc_b200 with extra nvcc defines (kernel experiments):  tools/build_alt.sh NAME -DFOO=1 ...
# -> aes_lac_2018_b200/lib/libctc_b200_NAME.so ; run with CTC_B200_LIB=<that path>
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/aes_lac_2018_b200/csrc; out=$root/aes_lac_2018_b200/build/alt_$name; mkdir -p $out
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas=-v $*"
pids=()
for g in 0 1 2 3 4 5; do nvcc $flags -DCTC_GROUP=$g -c $src/ctc_variants.cu -o $out/g$g.o > $out/g$g.log 2>&1 & pids+=($!); done
nvcc $flags -c $src/ctc_abi.cu -o $out/abi.o > $out/abi.log 2>&1 & pids+=($!)
nvcc $flags -c $src/ctc_head.cu -o $out/head.o > $out/head.log 2>&1 & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/aes_lac_2018_b200/lib/libctc_b200_$name.so $out/*.o
echo built $root/aes_lac_2018_b200/lib/libctc_b200_$name.so
