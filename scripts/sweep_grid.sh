for ctas in 37 148 296; do for chunk in 1048576 8388608 33554432; do
echo "ctas=$ctas chunk=$chunk: $(FSGPU_SWEEP_CTAS=$ctas FSGPU_SWEEP_CHUNK=$chunk python scripts/run_op.py q4rs 1000 4 2>&1 | tail -1)"
done; done
