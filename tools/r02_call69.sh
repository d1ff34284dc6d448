#!/bin/bash
# the clean rebuild of the library: smoke + a cross-section of the parity suite
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c1-300
timeout 300 python -m pytest tests -x -q -m gpu -k "sweep_parity or algorithm3 or run_chains or driver" 2>&1 | tail -n 2 | cut -c1-200
