for cfg in "8 8 4 1" "8 8 8 1" "8 8 16 1" "8 16 8 1" "8 16 16 1" "16 8 4 1" "16 8 8 1" "16 16 8 1" "8 8 2 1" "16 8 2 1"; do
  python profiles/one_config.py $cfg 10 125000 10000 float32 2>&1 | tail -1
done
