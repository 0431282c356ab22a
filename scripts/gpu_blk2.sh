#!/bin/bash
for g in 2 4 8 12; do echo "== scatter grid $g"; SBK_LIB_VARIANT=probe SBK_SCATTER_GRID=$g python scripts/dbg_blk.py 2>&1 | grep tensor_symmetric | tail -1 | cut -c1-300; done
