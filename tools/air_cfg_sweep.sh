#!/bin/bash
# GPU box: sweep the constraint-kernel launch configurations (block size x blocks/SM; inlined vs out-of-line multiply), print stage times
for call in 0 1; do for v in 1 2 3 4 5 6; do echo "DG_AIR_CALL=$call DG_AIR_CFG=$v"; DG_AIR_CALL=$call DG_AIR_CFG=$v python tools/stage_times.py 20 2 2>&1 | tail -1; done; done
