#!/bin/bash
# steady-state bench line of several env kinds
for spec in "HumanoidPyBulletEnv-v0 2048" "HumanoidPyBulletEnv-v0 4096" "HumanoidFlagrunPyBulletEnv-v0 2048" "AntMuJoCoEnv-v0 4096" "HumanoidMuJoCoEnv-v0 2048" "HalfCheetahMuJoCoEnv-v0 4096" "Walker2DMuJoCoEnv-v0 4096" "InvertedPendulumPyBulletEnv-v0 4096" "InvertedDoublePendulumPyBulletEnv-v0 4096" "ReacherPyBulletEnv-v0 4096" "AntPyBulletEnv-v0 65536"; do
  set -- $spec
  python bench.py --env $1 --envs $2 --steps 200 --warmup 20 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s %6d  value %.4g  ms %.4f  resident %.4g  e2e %.4g / %.4g  frac %s' % ('$1', $2, l['value'], l['ms_per_step'], l['value_l2_resident'], l['e2e']['value'], l['e2e']['value_staged_copies'], l['roofline']['frac']))"
done
