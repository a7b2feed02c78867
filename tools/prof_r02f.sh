# end-of-round-2 evidence: launch list of one window with the final defaults (selective PDL, early phone projection), ncu
# --set full of the fused residual-block kernel (opt-in, RVC_CBR=2) and of the tcgen05 GEMM inside a 32-window batched plan,
# memcheck of one window
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
NCU="ncu --set full --clock-control none --import-source on"
python tools/one_window.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_window.py > gpurun_out/ncu_window.log 2>&1
if [ -n "$FULL" ]; then
WINDOWS=2 GRAPH=0 RVC_CBR=2 $NCU -k regex:cbr_kernel -s 9 -c 3 -o gpurun_out/prof_cbr python tools/one_window.py > gpurun_out/p1.log 2>&1
NBS=32 SKIP_CHECK=1 $NCU -k regex:umma_gemm -s 40 -c 4 -o gpurun_out/prof_umma_b32 python tools/batch_check.py > gpurun_out/p9.log 2>&1
fi
WINDOWS=1 GRAPH=0 timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/one_window.py > gpurun_out/sanitizer_memcheck.log 2>&1
cat gpurun_out/ncu_window.log gpurun_out/sanitizer_memcheck.log
