#!/bin/bash
# debugging aid: the BSDP protein2genome golden command line with and without the specialised
# kernels, one trace line per Viterbi_calculate; prints the first differing calls
cd "$(dirname "$0")/../tests/golden/cli"
ARGS=$(python -c "import json; print(' '.join(json.load(open('commands.json'))['bsdp_protein2genome']))")
BIN=../../../integration/_build/exonerate_b200
EXONERATE_B200_TRACE=1 $BIN $ARGS > /tmp/t_int.out 2> /tmp/t_int.err
EXONERATE_B200_TRACE=1 C4B_GENERIC_JIT=1 C4B_JIT_SYSTOLIC=0 $BIN $ARGS > /tmp/t_jit.out 2> /tmp/t_jit.err
EXONERATE_B200_TRACE=1 C4B_GENERIC_JIT=1 $BIN $ARGS > /tmp/t_sys.out 2> /tmp/t_sys.err
wc -l /tmp/t_int.out /tmp/t_jit.out /tmp/t_sys.out /tmp/t_int.err /tmp/t_jit.err /tmp/t_sys.err
echo "== interpreter vs thread-per-row JIT"; diff <(grep b200-trace /tmp/t_int.err) <(grep b200-trace /tmp/t_jit.err) | head -12
echo "== interpreter vs systolic JIT"; diff <(grep b200-trace /tmp/t_int.err) <(grep b200-trace /tmp/t_sys.err) | head -12
grep -v b200-trace /tmp/t_sys.err | head -5
