"""Decode the scheduling control bits (stall count, yield, scoreboard barriers) of `cuobjdump -sass` output.
usage: cuobjdump -sass -fun NAME lib.so | python tools/sass_ctrl.py [grep-pattern]"""
import re
import sys

pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
lines = [l.rstrip('\n') for l in sys.stdin]
i = 0
while i < len(lines) - 1:
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', lines[i])
    m2 = re.search(r'^\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
    if m and m2:
        full = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
        stall, yld = (full >> 105) & 0xf, (full >> 109) & 1
        wbar, rbar, wait = (full >> 110) & 7, (full >> 113) & 7, (full >> 116) & 0x3f
        if pat is None or pat.search(m.group(2)):
            print('%s %-52s stall=%2d yield=%d wbar=%d rbar=%d wait=%s' % (m.group(1), m.group(2)[:52], stall, yld, wbar, rbar, format(wait, '06b')))
        i += 2
    else:
        i += 1
