for cfg in "0 0" "1 0" "3 0" "7 0" "15 0" "0 1" "1 1"; do set -- $cfg
  echo "=== MTB_TC_DEBUG=$1 MTB_DISABLE_PATCH=$2"; MTB_TC_DEBUG=$1 MTB_DISABLE_PATCH=$2 timeout 300 python scripts/op_profile.py --batch 128 --top 60 2>&1 | grep -E "total backbone|1\.1\.0\.block\.0 |1\.2\.1\.block\.0 |1\.3\.1\.block\.0 |1\.5\.1\.block\.0 |1\.5\.1\.block\.3 |1\.2\.1\.block\.1 " | cut -c1-125
done
