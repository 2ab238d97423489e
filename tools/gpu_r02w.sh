#!/bin/bash
bash tools/gpu_round2.sh r02w ncu ncuvjp
bash tools/gpu_profile_families.sh r02w
