cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" CHAINS=1 timeout 300 python tools/quick_ms.py > gpurun_out/t_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/t_$tag.log; }
run base RVC_NOP=1
run bw16 RVC_CHAIN_BW=16
run bw16_s64 RVC_CHAIN_BW=16 RVC_CHAIN_SIDE=64
run bw16_s96 RVC_CHAIN_BW=16 RVC_CHAIN_SIDE=96
run bw16_s148 RVC_CHAIN_BW=16 RVC_CHAIN_SIDE=148
run bw8_s96_sk2 RVC_CHAIN_BW=8 RVC_CHAIN_SIDE=96 RVC_CHAIN_SK=2500
run bw16_s96_m72 RVC_CHAIN_BW=16 RVC_CHAIN_SIDE=96 RVC_CHAIN_SIDE_MAXM=72 RVC_SC_LANE=0
grep -H -E "QUICK|rc=[^0]|^chain" gpurun_out/t_*.log
