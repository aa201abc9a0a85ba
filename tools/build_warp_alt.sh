#!/bin/bash
# Alternative warp-ladder table (kernel experiments):  tools/build_warp_alt.sh NAME "VW_(8,8,168),VW_(16,4,168)" [-DFOO=1 ...]
# -> aes_lac_2018_b200/lib/libctc_b200_NAME.so (all other objects come from the default build); run with CTC_B200_LIB=<path>
set -e
name=$1; table=$2; shift; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/aes_lac_2018_b200/csrc; out=$root/aes_lac_2018_b200/build/walt_$name; mkdir -p $out
echo "$table," > $out/table.inc
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas=-v $*"
for g in 6 7; do nvcc $flags -DCTC_GROUP=$g "-DCTC_WARP_TABLE_INC=\"$out/table.inc\"" -c $src/ctc_variants.cu -o $out/g$g.o > $out/g$g.log 2>&1 & done
wait
grep -E "error" $out/g6.log $out/g7.log && exit 1
objs=$(ls $root/aes_lac_2018_b200/build/*.o | grep -v -E "ctc_variants_g[67].o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/aes_lac_2018_b200/lib/libctc_b200_$name.so $objs $out/g6.o $out/g7.o
grep -E "Compiling|registers|spill" $out/g6.log | sed -e 's/ptxas info    : //' | paste - - - | sed -e "s/Compiling entry function '_ZN7ctcb20015ctc_warp_kernelI//" -e "s/EEvNS_11FusedParamsE' for 'sm_100a'//" -e "s/0 bytes stack frame, //"
echo built $root/aes_lac_2018_b200/lib/libctc_b200_$name.so
