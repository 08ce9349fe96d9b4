#!/bin/bash
# Evidence run for profiles/ (one GPU).  Each ncu command runs only after the same program exited 0 without ncu.
set -u
O=gpurun_out
T=${1:-r02n}
python tools/full_forward.py 16 128 2 > $O/p_full.log 2>&1 || { echo "full_forward failed"; exit 1; }
python tools/bench_attn.py 16 128 > $O/${T}_bench_attn.txt 2>&1 || { echo "bench_attn failed"; exit 1; }
python tools/bench_tail.py 16 128 > $O/${T}_bench_tail.txt 2>&1 || { echo "bench_tail failed"; exit 1; }
python tools/bench_dw.py 16 128 > $O/${T}_bench_dw.txt 2>&1 || { echo "bench_dw failed"; exit 1; }
python tools/bench_naf.py 16 512 > $O/${T}_bench_naf.txt 2>&1 || { echo "bench_naf failed"; exit 1; }
python tools/bench_small_convs.py 16 512 > $O/${T}_bench_small_convs.txt 2>&1 || { echo "bench_small_convs failed"; exit 1; }
# 1. launch list of one steady-state step at the bench shape (experts serialised for attribution): the last forward of
#    tools/full_forward.py sits between cudaProfilerStart / Stop, so every kernel of that step is listed (torch's too, if any)
FFB200_EXPERT_STREAMS=0 FFB200_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --profile-from-start off --csv --log-file $O/${T}_launches_B16_S128.csv python tools/full_forward.py 16 128 3 > $O/p_ncu1.log 2>&1
python tools/agg_launches.py $O/${T}_launches_B16_S128.csv > $O/${T}_full_model_B16_S128_launches.txt
# 2. --set full of 16 consecutive conv_gemm launches inside HAT blocks (DRAM traffic, tensor pipe, issue)
FFB200_EXPERT_STREAMS=0 FFB200_GRAPHS=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_gemm_tc -s 60 -c 16 -o $O/${T}_conv_gemm -f python tools/full_forward.py 16 128 1 > $O/p_ncu2.log 2>&1
# 3. --set full of the attention kernels (W-MSA on the 4-CTA kernel, OCAB, DAT 8x32): the 4th launch of each case
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:window_attention_tc4 -s 3 -c 1 -o $O/${T}_attn_wmsa -f python tools/bench_attn.py 16 128 > $O/p_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:ocab_attention -s 3 -c 1 -o $O/${T}_attn_ocab -f python tools/bench_attn.py 16 128 > $O/p_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:Geo<8" -s 3 -c 1 -o $O/${T}_attn_dat -f python tools/bench_attn.py 16 128 > $O/p_ncu5.log 2>&1
# 4. the fused HAT-block tail and the SimpleGate depthwise kernel
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:hab_tail -s 30 -c 1 -o $O/${T}_hab_tail -f python tools/bench_tail.py 16 128 > $O/p_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dwconv3x3_gate -s 3 -c 1 -o $O/${T}_dwgate -f python tools/bench_dw.py 16 128 > $O/p_ncu7.log 2>&1
# 5. the fused NAFBlock tail
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:naf_tail -s 6 -c 1 -o $O/${T}_naf_tail -f python tools/bench_naf.py 16 512 > $O/p_ncu8.log 2>&1
ls -la $O/*.ncu-rep
