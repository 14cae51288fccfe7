#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider -k "tcgen05_conv_leaky_pool or cnn_batch128 or vgg_small or training_matches" 2>&1 | tail -4
timeout 300 python scripts/steady_parts.py cnn 128 2>&1 | tail -2
