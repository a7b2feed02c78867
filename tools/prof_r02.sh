# round-2 evidence run: ncu launch list of one window, ncu --set full captures of the kernels VERDICT r1 named,
# compute-sanitizer memcheck + racecheck of one window with the persistent kernels on.  Outputs: gpurun_out/
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/one_window.py > /dev/null 2>&1   # builds the weight files
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_window.py > gpurun_out/ncu_window.log 2>&1
WINDOWS=2 GRAPH=0 $NCU -k regex:stft_mel_log -s 1 -c 1 -o gpurun_out/prof_stft python tools/one_window.py > gpurun_out/p1.log 2>&1
WINDOWS=2 GRAPH=0 $NCU -k regex:f0_decode -s 1 -c 1 -o gpurun_out/prof_f0decode python tools/one_window.py > gpurun_out/p2.log 2>&1
WINDOWS=2 GRAPH=0 $NCU -k regex:gemm_v2 -s 100 -c 3 -o gpurun_out/prof_gemmv2 python tools/one_window.py > gpurun_out/p3.log 2>&1
WINDOWS=2 GRAPH=0 $NCU -k regex:knn_scan -s 1 -c 1 -o gpurun_out/prof_knn768 python tools/one_window.py > gpurun_out/p4.log 2>&1
ONLY1M=1 $NCU -k regex:knn_umma_scan -s 2 -c 1 -o gpurun_out/prof_knn1m python tools/knn_bench.py > gpurun_out/p5.log 2>&1
WINDOWS=2 GRAPH=0 RVC_CVSTACK=1 $NCU -k regex:cvstack -s 1 -c 1 -o gpurun_out/prof_cvstack python tools/one_window.py > gpurun_out/p6.log 2>&1
WINDOWS=2 GRAPH=0 $NCU -k regex:chain_kernel -s 3 -c 1 -o gpurun_out/prof_chain python tools/one_window.py > gpurun_out/p7.log 2>&1
WINDOWS=2 GRAPH=0 $NCU -k regex:umma_gemm -s 131 -c 3 -o gpurun_out/prof_umma python tools/one_window.py > gpurun_out/p8.log 2>&1
$NCU -k regex:umma_gemm -s 40 -c 4 -o gpurun_out/prof_umma_b32 NBS=32 SKIP_CHECK=1 python tools/batch_check.py > gpurun_out/p9.log 2>&1
WINDOWS=1 GRAPH=0 RVC_CVSTACK=1 timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/one_window.py > gpurun_out/sanitizer_memcheck.log 2>&1
WINDOWS=1 GRAPH=0 RVC_CVSTACK=1 timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/one_window.py > gpurun_out/sanitizer_racecheck.log 2>&1
ls -la gpurun_out | tail -30
tail -3 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log
