#!/bin/bash
# Round-2 evidence run on the GPU box: ncu --set full captures of one launch per kernel class (exported as raw CSV, the
# .ncu-rep files are dropped to stay under the transfer limit), compute-sanitizer memcheck / racecheck on small tests.
# Usage (from the repo root, under gpurun): bash tools/r2_profile.sh [ncu|sanitize|all]
set -u
what=${1:-all}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --cache-control none --import-source off --kernel-name-base demangled"
cap() {  # name, kernel regex, launch-skip, extra args of profile_classes.py
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 300 $NCU -k "regex:$regex" -s $skip -c 1 -f -o $out/$name python tools/profile_classes.py "$@" > $out/$name.log 2>&1
  if [ -f $out/$name.ncu-rep ]; then
    ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.csv 2>/dev/null
    rm -f $out/$name.ncu-rep
    echo "captured $name: $(wc -c < $out/$name.csv) bytes"
  else
    echo "FAILED $name"; tail -3 $out/$name.log
  fi
}
if [ "$what" = ncu ] || [ "$what" = all ]; then
 if [ -z "${ONLY_FAILED:-}" ]; then
  cap r2_ncu_chain_k0       chain_tc                  13 --max-length 4      # K0 of the second decode step
  cap r2_ncu_chain_kb       chain_tc                  14 --max-length 4      # KB[0]
  cap r2_ncu_chain_ka       chain_tc                  15 --max-length 4      # KA[0]
  cap r2_ncu_rmsnorm        rmsnorm_kernel            2  --max-length 2
  cap r2_ncu_select_token   select_token_kernel       1  --max-length 4
  cap r2_ncu_fold_split     fold_split_kernel         1  --max-length 2 --mel-segments 256
  cap r2_ncu_mel_band_log   mel_band_log_kernel       1  --max-length 2 --mel-segments 256
 fi
  cap r2_ncu_dft_gemm       "gemm_tc_kernel.*EpiDftRe" 1 --max-length 2 --mel-segments 256
  [ -z "${ONLY_FAILED:-}" ] && cap r2_ncu_enc_attn_tc    enc_attn_tc_kernel        1  --max-length 2
  cap r2_ncu_gemm_tc2_qkv   "gemm_tc2_kernel.*256.*EpiStore" 1 --max-length 2
  cap r2_ncu_gemm_tc2_res   "gemm_tc2_kernel.*192.*EpiResidual" 2  --max-length 2
  cap r2_ncu_dec_self_attn  "decode_attn_kernel<__nv_bfloat16, .bool.1"  1800 --max-length 304   # layer 0 of decode step 300
  cap r2_ncu_dec_cross_attn "decode_attn_kernel<__nv_bfloat16, .bool.0" 60   --max-length 12
  # launch list (device time per launch) of the bench command, two full decode steps around step 60 (ncu intercepts every
  # launch, ~40 ms each even when skipped, so step 500 of the full-length run is out of reach; the whole-run shares come
  # from tools/kernel_breakdown.py = CUPTI activity records, no replay)
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2050 -c 60 --csv --log-file $out/r2_launches_step60.csv \
      python bench.py --steps 1 --warmup 0 --max-length 64 --no-fp32 --no-extra --no-cpu-baseline --no-roofline > $out/r2_launches_step60.log 2>&1
  echo "launch list rc=$? $(wc -l < $out/r2_launches_step60.csv) lines"
fi
if [ "$what" = attn ]; then  # decode attention after the packed-arithmetic change
  cap r2_ncu_dec_self_attn  "decode_attn_kernel<__nv_bfloat16, .bool.1"  1800 --max-length 304   # layer 0 of decode step 300
  cap r2_ncu_dec_cross_attn "decode_attn_kernel<__nv_bfloat16, .bool.0" 60   --max-length 12
  exit 0
fi
if [ "$what" = late ]; then  # kernels that changed after the first evidence run of the round
  cap r2_ncu_enc_attn_tc    enc_attn_tc_kernel        1  --max-length 2
  cap r2_ncu_gemm_tc2_qkv   "gemm_tc2_kernel.*256.*EpiHeadMajorQKV" 1 --max-length 2
  cap r2_ncu_cross_kv       "gemm_tc2_kernel.*EpiHeadMajorKV" 0 --max-length 2
  cap r2_ncu_select_token   select_token_kernel       1  --max-length 4
  cap r2_ncu_seq_attn_tc_self  seq_attn_tc_kernel     2  --max-length 2 --clips 8 --teacher-forced 32   # decoder layer 1, causal
  cap r2_ncu_seq_attn_tc_cross seq_attn_tc_kernel     3  --max-length 2 --clips 8 --teacher-forced 32   # decoder layer 1, cross
  cap r2_ncu_f32_split_gemm "gemm_tc_kernel<.int.128, .int.2, .int.3, m2m::EpiStore" 2 --max-length 2 --precision fp32 --clips 64
  cap r2_ncu_f32_split3     split3_kernel             2  --max-length 2 --precision fp32 --clips 64
  cap r2_ncu_f32_dec_self_attn "decode_attn_kernel<float, .bool.1" 600 --max-length 104 --precision fp32
fi
if [ "$what" = sanitize ] || [ "$what" = all ] || [ "$what" = late ]; then
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
      -k "gemm_paths or logmel_matches or bf16_batch_sizes or encoder_bf16 or conditioning or eos_pad or fp32_split_product or frontend_is_fp32" > $out/r2_sanitizer_memcheck.log 2>&1
  echo "memcheck rc=$?"; tail -4 $out/r2_sanitizer_memcheck.log
  [ "$what" = late ] && exit 0
  timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
      -k "bf16_batch_sizes and (1 or 129)" > $out/r2_sanitizer_racecheck_chain.log 2>&1
  echo "racecheck(chain) rc=$?"; tail -4 $out/r2_sanitizer_racecheck_chain.log
  timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -p no:cacheprovider \
      -k "gemm_paths and (4 or 5)" > $out/r2_sanitizer_racecheck_gemm2.log 2>&1
  echo "racecheck(gemm_tc2) rc=$?"; tail -4 $out/r2_sanitizer_racecheck_gemm2.log
fi
