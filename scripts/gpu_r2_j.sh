#!/bin/bash
python scripts/dbg_cp.py 2>&1 | tail -12
