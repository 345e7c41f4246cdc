for st in 2 3 4 5; do AMB_SWEEP_CHILD=1 AMB_PASSES=1 AMB_DEBUG_SINGLE=2 AMB_CTA2=1 AMB_STAGES=$st python profiles/engine_sweep.py 100000 512 2>&1 | grep passes; done
for g in 74 36; do AMB_SWEEP_CHILD=1 AMB_PASSES=1 AMB_DEBUG_SINGLE=2 AMB_CTA2=1 AMB_GRID=$g python profiles/engine_sweep.py 100000 512 2>&1 | grep passes; done
AMB_SWEEP_CHILD=1 AMB_PASSES=1 AMB_DEBUG_SINGLE=2 AMB_CTA2=1 python profiles/engine_sweep.py 100000 256 2>&1 | grep passes
AMB_SWEEP_CHILD=1 AMB_PASSES=1 AMB_DEBUG_SINGLE=1 AMB_CTA2=0 python profiles/engine_sweep.py 100000 256 2>&1 | grep passes
