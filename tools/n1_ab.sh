# N = 1 A/B of environment toggles on the configs[1] training step (graph replay): tag VAR=value ...
run() { tag=$1; shift; env "$@" python bench.py --no-sliding-window --no-ranking --no-cpu-baseline --no-dp128 --no-augment 2>gpurun_out/ab_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step'],4), round(d['value'],1))"; }
run default A=1
run read_act B200_INBWD_ACT=1
run default2 A=1
