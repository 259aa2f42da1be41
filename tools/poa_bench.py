#!/usr/bin/env python
"""The insertion-alignment kernel on the alignment problems of configs[4]-derived subgroups (bench.py's `poa` block on
its own).   usage: poa_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rambl_b200 import api
print(bench.poa_block(api))
