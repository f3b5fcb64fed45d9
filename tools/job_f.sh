python tools/prof_facade.py 2>&1 | tail -45
