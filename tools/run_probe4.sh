#!/bin/bash
# runs the selected groups of tools/umma_probe4, each in its own process
for a in "3 0" "3 1" "3 2"; do
  timeout 60 tools/bin/umma_probe4 $a 2>&1
  echo "-- group $a rc=$?"
done
