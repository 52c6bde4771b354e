#!/bin/bash
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --timeout-method thread 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
