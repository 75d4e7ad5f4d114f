#!/bin/bash
# 2-GPU experiment: does the overlapped gradient all-reduce hide under the backward?  NCCL channel count x reserved SMs
mkdir -p gpurun_out
for cfg in "default 16" "8 16" "4 16" "8 32" "16 32" "default 0"; do
  set -- $cfg
  if [ "$1" = "default" ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$1; fi
  export RLCA_RESERVED_SMS=$2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus 2 --steps 200 --warmup 20 --e2e-steps 5 2>> gpurun_out/r2m_dp.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)['learner_dp']; d['nchannels'] = '$1'; d['reserved_sms'] = $2; print(json.dumps(d))
" >> gpurun_out/r2m_dp.jsonl
done
cat gpurun_out/r2m_dp.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['nchannels'], d['reserved_sms'], 'overlap', round(d['step_us'],1), 'serial', round(d['step_serial_allreduce_us'],1), 'none', round(d['step_without_allreduce_us'],1), 'ar', round(d['allreduce_alone_us'],1))
"
tail -3 gpurun_out/r2m_dp.err
