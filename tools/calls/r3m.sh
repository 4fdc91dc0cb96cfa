timeout 300 python tools/dp_debug3.py 2>&1 | grep -v Warning | tail -5 | cut -c1-1500
