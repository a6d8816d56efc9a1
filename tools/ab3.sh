for cfg in "$@"; do echo "== $cfg"; env $(echo $cfg | tr ',' ' ') python tools/timeline_probe.py 2>&1 | tail -1 | cut -c1-330; done
