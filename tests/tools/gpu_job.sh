#!/bin/bash
# the round's full GPU pass: parity tests + smoke + bench lines, then the ncu evidence and the density sweep
cd "$(dirname "$0")/../.."
bash tests/tools/gpu_job_final.sh
bash tests/tools/gpu_job_evidence.sh
