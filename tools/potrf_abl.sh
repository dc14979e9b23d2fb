for v in 0 1 2 3; do echo variant $v; BO_POTRF_DBG=$v python tools/potrf_abl.py 2>&1 | tail -2; done
