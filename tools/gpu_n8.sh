run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/dp_step_time.py 2>&1 | grep "us/step"; }
run A=1
run B200CLIP_BWD_SPLITS=4
run B200CLIP_NCCL_PRIO=0
run B200CLIP_TWO_STREAM_ROWS=0
run NCCL_PROTO=LL128
bash tools/gpu_n.sh 8 r2o 2>&1 | grep -v Traceback | head -5
