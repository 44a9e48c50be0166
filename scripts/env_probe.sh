#!/bin/bash
# dev: the 16K^2 terrain build under a list of environment settings, e.g.  scripts/env_probe.sh "CPVS_RANK_ONEPASS=0" "CPVS_RANK_ONEPASS=1"
for setting in "$@"; do echo "== $setting"; env $setting python scripts/one_build.py 16384 terrain 5 2>&1 | tail -2; done
