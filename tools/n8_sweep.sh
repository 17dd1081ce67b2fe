run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-ranking --no-dp128 --no-augment --no-dp-check $EXTRA 2>gpurun_out/n8_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', d['ms_per_step'], d['value'], (d.get('sliding_window') or {}).get('ms_per_volume'))"; }
EXTRA=--no-sliding-window
run default A=1
run ctas8 NCCL_MAX_CTAS=8
run ctas4 NCCL_MAX_CTAS=4
run groups4 B200_GRAD_GROUPS=4
run groups13 B200_GRAD_GROUPS=13
EXTRA=
run swtiming B200_SW_TIMING=1
grep "sliding window rank" gpurun_out/n8_swtiming.err | tail -8
