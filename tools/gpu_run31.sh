#!/bin/bash
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread 2>&1 | tail -8
