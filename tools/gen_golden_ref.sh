#!/bin/bash
# Generates golden vectors by running the UNMODIFIED reference binary (oracle/_ref/cuIsing_ref,
# built from /root/reference/optimized by oracle/Makefile) on the GPU box.  Outputs go to
# gpurun_out/golden/; tools/pack_golden.py turns them into the fixtures under tests/golden/.
set -u
REF=$PWD/oracle/_ref/cuIsing_ref
OUT=$PWD/gpurun_out/golden
if [ "${1:-all}" = "extra" ]; then OUT=$PWD/gpurun_out/golden_extra; fi
rm -rf $OUT; mkdir -p $OUT
run() { # name, args...
  name=$1; shift
  mkdir -p $OUT/$name; (cd $OUT/$name && timeout 600 $REF "$@" > stdout.txt 2> stderr.txt; echo "$@" > args.txt)
  grep -E "magnetization|Kernel execution" $OUT/$name/stdout.txt | tail -3
}
if [ "${1:-all}" = "extra" ]; then
run corr_4096       -x 4096 -y 1024 -n 32 -p 16 -t 2.269185 -c
run corr_8192_T2    -x 8192 -y 512 -n 16 -p 8 -t 2.0 -c -s 987654321
run bonds_sub       -x 4096 -y 64 -n 4 -p 2 -t 1.5 -J 0.3 --xsl 2048 --ysl 32 -o
tar czf $PWD/gpurun_out/golden_extra.tgz -C $PWD/gpurun_out golden
exit 0
fi
run small_T1.5      -x 2048 -y 16 -n 4 -p 1 -t 1.5 -o
run small_Tc        -x 2048 -y 16 -n 4 -p 1 -t 2.269185 -o
run rag_Tc          -x 6144 -y 48 -n 3 -p 1 -t 2.269185 -o
run sub_T1.5        -x 4096 -y 64 -n 8 -p 4 -t 1.5 --xsl 2048 --ysl 16 -o
run alpha_default   -x 2048 -y 32 -n 6 -p 2 -o
run seed_T2         -x 4096 -y 32 -n 5 -p 5 -t 2.0 -s 1311768467294899695 -o
run mid_T1.5        -x 4096 -y 4096 -n 16 -p 4 -t 1.5
run mid_Tc          -x 4096 -y 4096 -n 16 -p 4 -t 2.269185
run big_Tc          -x 65536 -y 65536 -n 128 -p 16 -t 2.269185
run big_T1.5        -x 65536 -y 65536 -n 128 -p 16 -t 1.5
run big_sub_T1.5    -x 65536 -y 65536 -n 32 -p 16 -t 1.5 --xsl 2048 --ysl 2048
run ramp            -x 2048 -y 64 -n 12 -p 3 -t 3.0 -u -0.25,3 -o
run bonds_J0.25     -x 2048 -y 32 -n 4 -p 1 -t 1.5 -J 0.25 -o
run bonds_J0.5_Tc   -x 4096 -y 64 -n 6 -p 3 -t 2.269185 -J 0.5 -o
run corr_Tc         -x 2048 -y 2048 -n 64 -p 16 -t 2.269185 -c
run corr_sub        -x 4096 -y 4096 -n 32 -p 16 -t 2.0 -c --xsl 2048 --ysl 2048
run expser          -x 2048 -y 64 -n 700 -e -t 2.269185
run magn_stop       -x 2048 -y 2048 -n 400 -p 10 -t 1.5 -m 0.5
du -sh $OUT; ls $OUT
# keep big directories small: only stdout for the big runs
find $OUT -name 'lattice_*' -size +3M -delete
tar czf $PWD/gpurun_out/golden.tgz -C $PWD/gpurun_out golden
ls -la $PWD/gpurun_out/golden.tgz
