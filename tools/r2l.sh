#!/bin/bash
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
