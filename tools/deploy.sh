#!/bin/bash
# Working copy .dev/ -> repository root (the tree gpurun snapshots).  Edits are made in .dev/ so that a
# queued gpurun call never snapshots a half-edited tree; run this before every gpurun call and commit.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
case "$here" in */.dev/tools) root="${here%/.dev/tools}";; *) root="${here%/tools}";; esac
cd "$root"
tar -C .dev --exclude='__pycache__' --exclude=.pytest_cache --exclude=.hypothesis --exclude=./gpurun_out -cf - . | tar -xf - -C .
