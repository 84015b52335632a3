#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python tools/diag_parity.py isotherm_gp 2>&1 | tail -40 ) > gpurun_out/diag_parity.log
cat gpurun_out/diag_parity.log
