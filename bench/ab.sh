#!/bin/bash
# developer A/B harness: run the quick benches with each library named on the command line
# usage: bench/ab.sh libA.so libB.so ...   (paths relative to glome_b200/_build/)
cd "$(dirname "$0")/.."
for lib in "$@"; do
  export GLOME_LIB=$PWD/glome_b200/_build/$lib
  echo "=== $lib"
  python bench/quick.py 2 1000000 1920 1080 0
  python bench/quick.py 2 1000000 720 480 1
  python bench/quick_shard.py 8
  python bench/quick_shard.py 4
  python bench/quick.py 3 2000000 3840 2160 1
done
