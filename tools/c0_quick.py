"""Runs bench.py's configs[0] leg alone (the real binaries on an ETI file) and prints its JSON."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dabmod_loader  # noqa: E402

dabmod_loader.load()
import bench  # noqa: E402

print(json.dumps(bench.measure_binary(), indent=1))
