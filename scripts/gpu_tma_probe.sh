timeout 120 ./scripts/dev/tma_probe
