# does L2 residency (small sub-batches) help? per-op profile at B=32/64/128, plus SE-fusion variant at 64
mkdir -p gpurun_out
for b in 32 64 128; do
  echo "=== op profile B=$b"; timeout 300 python scripts/op_profile.py --batch $b --top 30 2>&1 | cut -c1-200 | tee gpurun_out/op_profile_b$b.txt | head -34
done
echo "=== op profile B=64 MTB_FUSE_SE=1"; MTB_FUSE_SE=1 timeout 300 python scripts/op_profile.py --batch 64 --top 30 2>&1 | cut -c1-200 | tee gpurun_out/op_profile_b64_fuse.txt | head -34
echo "=== op profile B=256 MTB_FUSE_SE=1"; MTB_FUSE_SE=1 timeout 300 python scripts/op_profile.py --batch 256 --top 30 2>&1 | cut -c1-200 | tee gpurun_out/op_profile_b256_fuse.txt | head -34
