for v in 0 1 2 3 4 5 6; do echo "variant $v: $(FDFD_APPLY_VARIANT=$v python tools/prof_apply.py 4096 200 | tail -1)"; done
