#!/bin/bash
# Alternative fp32 warp-ladder table (kernel experiments):  tools/build_warp32_alt.sh NAME "VF_(8,8,128),VF_(16,8,168)" [-DFOO=1 ...]
# -> aes_lac_2018_b200/lib/libctc_b200_NAME.so (all other objects come from the default build); run with CTC_B200_LIB=<path>
set -e
name=$1; table=$2; shift; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/aes_lac_2018_b200/csrc; out=$root/aes_lac_2018_b200/build/w32alt_$name; mkdir -p $out
echo "$table," > $out/table.inc
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas=-v $*"
for g in 8 9; do nvcc $flags -DCTC_GROUP=$g "-DCTC_WARP32_TABLE_INC=\"$out/table.inc\"" -c $src/ctc_variants.cu -o $out/g$g.o > $out/g$g.log 2>&1 & done
wait
grep -E "error" $out/g8.log $out/g9.log && exit 1
objs=$(ls $root/aes_lac_2018_b200/build/*.o | grep -v -E "ctc_variants_g[89].o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/aes_lac_2018_b200/lib/libctc_b200_$name.so $objs $out/g8.o $out/g9.o
grep -E "Compiling|registers|spill" $out/g8.log | sed -e 's/ptxas info    : //' | paste - - - | sed -e "s/Compiling entry function '_ZN7ctcb20017ctc_warp32_kernelI//" -e "s/EEvNS_11FusedParamsE' for 'sm_100a'//" -e "s/0 bytes stack frame, //" -e "s/Function properties for [^ ]*//"
echo built $root/aes_lac_2018_b200/lib/libctc_b200_$name.so
