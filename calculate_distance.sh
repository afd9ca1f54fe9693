#!/bin/bash
# Drop-in for the reference's calculate_distance.sh (xvfb-run python3 $PWD/calculate_distance.py $@):
# no X server is needed here because no PNG is rendered (ete3 is not installed).
exec python3 "$(dirname "$0")/calculate_distance.py" "$@"
