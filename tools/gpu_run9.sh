#!/bin/bash
TWXI_LIB=$PWD/topowx_b200/libtwxi_prof.so timeout 600 python tools/ked_prof.py 2>&1 | tail -22
