python tools/profile_stage1.py 1024 cl 2>&1 | grep -v Warn | head -48 | cut -c1-200
