"""Generates the host-CLI fixtures from the reference tree (run in the dev container, where /root/reference exists):
  column_docs.txt   "NAME: text" lines recovered from web/views/_plaac_headers.haml, the reference's own generated
                    golden of plaac.java's -d output (cli/build_docs.py turns `## NAME: text` lines into that HAML)
  column_names.txt  the 38 summary columns in order
  bg_freqs_HUMAN.txt  copy of the reference DATA file web/bg_freqs/bg_freqs_HUMAN.txt (input for -B / config 3)
"""
import html
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

lines = open(os.path.join(REF, "web/views/_plaac_headers.haml")).read().split("\n")
docs, names = [], []
i = 0
while i < len(lines):
    if lines[i].strip() == "%li":
        name = lines[i + 1].strip().replace("%strong ", "")
        text = html.unescape(lines[i + 2].strip())
        docs.append(f"{name}: {text}")
        names.append(name)
        i += 3
    else:
        i += 1
open(os.path.join(HERE, "column_docs.txt"), "w").write("\n".join(docs) + "\n")
open(os.path.join(HERE, "column_names.txt"), "w").write("\n".join(names) + "\n")
shutil.copyfile(os.path.join(REF, "web/bg_freqs/bg_freqs_HUMAN.txt"), os.path.join(HERE, "bg_freqs_HUMAN.txt"))
print(len(docs), "columns")
