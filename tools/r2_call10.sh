# round 2, GPU call 10: full GPU suite (measured outlier counts written), then the evidence run of the default bench
export PBRS_WRITE_OUTLIERS=1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
cp profiles/parity_outliers.json profiles/parity_outliers_samples.json gpurun_out/ 2>/dev/null
unset PBRS_WRITE_OUTLIERS
bash tools/final_measure.sh 2>&1 | tail -30
