cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python tools/quick_ms.py > gpurun_out/s_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/s_$tag.log; }
run base RVC_NOP=1
run cv32 RVC_CV_WANT=32
run cv48 RVC_CV_WANT=48
run cv64 RVC_CV_WANT=64
run cv80 RVC_CV_WANT=80
run cv48_s64 RVC_CV_WANT=48 RVC_CHAIN_SIDE=64
run stack48 RVC_CVSTACK=1 RVC_CVSTACK_G=48
run stack64_cv48 RVC_CVSTACK=1 RVC_CV_WANT=48
grep -H -E "QUICK|rc=[^0]" gpurun_out/s_*.log
