cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" CHAINS=1 timeout 300 python tools/quick_ms.py > gpurun_out/r_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/r_$tag.log; }
run base RVC_NOP=1
run wpre RVC_CHAIN_WPRE=1
run poll RVC_CHAIN_POLL=1
run both RVC_CHAIN_WPRE=1 RVC_CHAIN_POLL=1
run base2 RVC_NOP=2
grep -H -E "QUICK|rc=[^0]|^chain" gpurun_out/r_*.log
