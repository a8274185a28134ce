import os, sys
sys.argv = [sys.argv[0], "200000"]
src = open('/root/repo/scripts/c5_probe.py').read().replace('for nb in (256, 512, 1024, 2048):', 'for nb in (1024,):').replace('for cl in ("auto", "1", "2", "4", "8"):', 'for cl in ("auto",):')
exec(compile(src, 'c5_probe', 'exec'))
