set -x
# A/B the v2 kernel variants for correctness (tiny + full) -- each under a timeout so a deadlock cannot eat the box
for mode in "1 1" "2 1" "1 2"; do
  set -- $mode
  export Q3TTS_TC_HALO=$1 Q3TTS_TC_CLUSTER=$2
  echo "=== HALO=$1 CLUSTER=$2"
  timeout 120 python tests/tools/debug_stages.py tiny fp16 2 9 2>&1 | grep -E "init_conv|block1|block3|pcm|rror"
  timeout 120 python tests/tools/debug_stages.py full fp16 2 20 2>&1 | grep -E "pre_conv|pre_transformer|upsample1|init_conv|block0|block1|block2|block3|pcm|rror"
done
