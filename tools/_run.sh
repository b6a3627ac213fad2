timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -3
for i in 1 2; do
echo "ksplit : $(timeout 300 python tools/train_profile.py bf16 1 32 2>&1 | tail -1)"
echo "nosplit: $(DFF_B200_WGRAD_NO_KSPLIT=1 timeout 300 python tools/train_profile.py bf16 1 32 2>&1 | tail -1)"
done
echo "ksplit B=4 : $(timeout 300 python tools/train_profile.py bf16 1 4 2>&1 | tail -1)"
echo "nosplit B=4: $(DFF_B200_WGRAD_NO_KSPLIT=1 timeout 300 python tools/train_profile.py bf16 1 4 2>&1 | tail -1)"
