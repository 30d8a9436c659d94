import re


def string2symbols(s):
    out = []
    for sym, count in re.findall(r'([A-Z][a-z]?)(\d*)', s):
        out.extend([sym] * (int(count) if count else 1))
    return out
