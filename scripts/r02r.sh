#!/bin/bash
# session R: SDM with programmatic dependent launch (tests, step times, stamps) + host-query pipeline with the ramped block schedule
bash scripts/r02n.sh
bash scripts/r02q.sh
