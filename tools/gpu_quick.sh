#!/bin/bash
# quick validation: GPU tests + stage timing of the benchmark tile
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
