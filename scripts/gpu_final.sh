# end-of-session evidence: GPU tests, then scripts/gpu_evidence.sh (bench lines of both arms, the other workloads, launch list, ncu captures)
T=${1:-final}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -15 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
bash scripts/gpu_evidence.sh $T
