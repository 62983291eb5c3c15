#!/bin/bash
# Freeze the current working tree (sources + built .so) under gpurun_pin/<name>/ so that a queued gpurun call
# runs exactly this state even if the main tree is edited while the call waits for a GPU slot.
# usage: scripts/pin_tree.sh <name>
set -e
cd "$(dirname "$0")/.."
dst=gpurun_pin/$1
rm -rf "$dst"; mkdir -p "$dst"
for f in kan-tts_b200 kantts_b200 oracle include tests scripts bench.py __graft_entry__.py MEASURED_PEAKS.json BASELINE.json; do
  [ -e "$f" ] && cp -a "$f" "$dst/"
done
mkdir -p "$dst/profiles"; cp -a profiles/*.json "$dst/profiles/" 2>/dev/null || true
echo "pinned $(du -sh $dst | cut -f1) -> $dst"
