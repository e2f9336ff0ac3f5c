"""Print the per-span timeline the library dumps with KS_TIMELINE=<path> (phase, stream, start_ms, end_ms)."""
import json, sys
names = ["featurize", "gramG", "allreduce", "solve", "update/AtR", "other"]
tl = json.load(open(sys.argv[1]))
lo = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hi = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
for ph, st, a, b in sorted(tl, key=lambda x: x[2]):
    if a >= lo and a <= hi:
        print(f"S{st} {names[ph]:11s} {a:9.3f} -> {b:9.3f}  ({b - a:7.3f} ms)")
