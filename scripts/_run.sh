scripts/gpu_ab.sh r02e "X=1|--batch 16" "X=1|--batch 24" "X=1|--batch 32" "X=1|--batch 48"
