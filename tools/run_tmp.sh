for t in 1 2 3; do DSRCGPU_TAIL_SPLIT=$t python tools/e2e_timeline.py 50000000 0 2>&1 | head -1; done
DSRCGPU_TAIL_SPLIT=2 DSRCGPU_SLOTS=4 python tools/e2e_timeline.py 50000000 0 2>&1 | head -1
DSRCGPU_TAIL_SPLIT=4 DSRCGPU_SLOTS=4 python tools/e2e_timeline.py 50000000 0 2>&1 | head -1
