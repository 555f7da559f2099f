for cfg in 0 1 3; do
  echo "=== MTB_TC_DEBUG=$cfg"; MTB_TC_DEBUG=$cfg timeout 300 python scripts/op_profile.py --batch 128 --top 60 2>&1 | grep -E "total backbone|1\.1\.0\.block\.0 |1\.2\.1\.block\.0 |1\.3\.1\.block\.0 |1\.5\.1\.block\.0 |1\.5\.1\.block\.3 |1\.2\.1\.block\.1 |1\.6\.1\.block\.0 |1\.6\.1\.block\.3 " | cut -c1-125
done
