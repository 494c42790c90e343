#!/usr/bin/env python
"""Build step: the device headers a run-time compiled density is built against, as byte
arrays for walnuts_b200/csrc/user_density.cu.  usage: embed_headers.py OUT.inc HEADER..."""
import sys
from pathlib import Path

out, headers = sys.argv[1], [Path(h) for h in sys.argv[2:]]
lines = []
for i, h in enumerate(headers):
    data = h.read_bytes() + b"\0"
    body = ",".join(str(b) for b in data)
    lines.append(f"static const unsigned char kEmb{i}[] = {{{body}}};")
names = ", ".join(f'"{h.name}"' for h in headers)
datas = ", ".join(f"reinterpret_cast<const char*>(kEmb{i})" for i in range(len(headers)))
lines.append(f"static const char* const kEmbeddedNames[] = {{{names}}};")
lines.append(f"static const char* const kEmbeddedData[] = {{{datas}}};")
lines.append(f"static const int kEmbeddedCount = {len(headers)};")
Path(out).write_text("\n".join(lines) + "\n")
