#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_cpp_binding.py -m gpu -x -q 2>&1 | tail -40
