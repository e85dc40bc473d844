#!/bin/bash
# same-box A/B of megab_decode_kernel variants: build/v_<name>/megab.o (compiled with -D switches) are linked here against
# the in-tree objects; build/wt_prev_so/libfsb.so is the library of the previous commit.
mkdir -p gpurun_out
C=fish_speech_rs_b200/csrc
OTHERS="$C/fsb_common.o $C/fsb_lm.o $C/fsb_lm_mega_bf16.o $C/fsb_lm_mega_f32.o $C/fsb_lm_mega1_bf16.o $C/fsb_lm_mega1_f32.o $C/fsb_tc_gemm.o $C/fsb_tc_conv.o $C/fsb_codec.o"
for d in build/v_*; do /usr/local/cuda/bin/nvcc -shared -o $d/libfsb.so $OTHERS $d/megab.o -lcudart 2>/dev/null; done
run_tests() { timeout -s KILL 400 python -m pytest tests/test_lm_gpu.py -x -q -m gpu -k "wide or batch or session or snapshot or ragged" 2>&1 | tail -2; }
[ -z "$AB_SKIP_CUR_TESTS" ] && { echo "== tests cur"; run_tests; }
for t in ${AB_TEST_VARIANTS:-all}; do echo "== tests $t"; FSB_LIB=$PWD/build/v_$t/libfsb.so run_tests; done
for v in $( [ -f build/wt_prev_so/libfsb.so ] && echo prev ) cur $(ls build | grep '^v_' | sed 's/^v_//'); do
  if [ $v = prev ]; then export FSB_LIB=$PWD/build/wt_prev_so/libfsb.so; elif [ $v = cur ]; then unset FSB_LIB; else export FSB_LIB=$PWD/build/v_$v/libfsb.so; fi
  for cfg in cfg5 cfg3; do
  timeout -s KILL 300 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/ab_${v}_${cfg}.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v $cfg', round(d['value']), {k:round(x,1) for k,x in d['breakdown_ms_per_step'].items()})"
  done
done
for v in ${AB_TIMER_VARIANTS:-cur all}; do
if [ $v = cur ]; then unset FSB_LIB; else export FSB_LIB=$PWD/build/v_$v/libfsb.so; fi
echo "== timers $v"
FSB_MEGA_TIMERS=1 timeout -s KILL 300 python bench.py --config cfg5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "\[megab" | tail -5
done
