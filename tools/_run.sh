for b in 4 8 16 32; do timeout 300 python tools/train_profile.py bf16 1 $b 2>&1 | tail -1; done
