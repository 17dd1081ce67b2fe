# N = 2 (or $NG): training-only bench legs with / without the deferred conv weight gradients
NG=${NG:-2}
run() { tag=$1; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 5 --no-ranking --no-dp128 --no-augment --no-dp-check --no-sliding-window --no-cpu-baseline 2>gpurun_out/defer_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', d['ms_per_step'], d['value'])"; }
run defer A=1
run inplace B200_NO_DEFER_WGRAD=1
run defer2 A=1
