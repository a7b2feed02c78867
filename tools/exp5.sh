cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" CHAINS=1 timeout 300 python tools/quick_ms.py > gpurun_out/u_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/u_$tag.log; }
run base RVC_NOP=1
run wide RVC_V2_NARROW=0
run side48 RVC_CHAIN_SIDE=48
run side64 RVC_CHAIN_SIDE=64
run side48m20 RVC_CHAIN_SIDE=48 RVC_CHAIN_SIDE_MAXM=20
run stack RVC_CVSTACK=1
run stack_s64 RVC_CVSTACK=1 RVC_CHAIN_SIDE=64
run stack_s84m20 RVC_CVSTACK=1 RVC_CHAIN_SIDE=84 RVC_CHAIN_SIDE_MAXM=20
run pdl RVC_PDL=1
grep -H -E "QUICK|rc=[^0]|^chain 0" gpurun_out/u_*.log
