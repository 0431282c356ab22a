#!/bin/bash
SBK_LIB_VARIANT=probe SBK_SYM_DEBUG=1 timeout 300 python scripts/dbg_sym_feat.py 450000 2>&1 | cut -c1-900
