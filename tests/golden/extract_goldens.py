#!/usr/bin/env python3
"""Extract the seed-42 golden vectors held by the reference's own inline tests.

Run in the build container (where /root/reference exists):

    python tests/golden/extract_goldens.py

It parses the Rust literals inside the `#[cfg(test)]` modules listed in
SURVEY.md Appendix C and writes them, untouched, to
tests/golden/reference_goldens.json.  Only expected VALUES are taken from the
reference (they are the known-answer tests for the un-vendored rand /
rand_distr / bio crates); no reference code is copied.  The JSON file is what
travels to the GPU box; nothing under tests/ reads /root/reference at run time.
"""
import json
import os
import re
import sys

REF = os.environ.get("RUSTYREAD_REFERENCE", "/root/reference")


class P:
    """Tiny recursive-descent parser for the Rust literal subset used in the tests."""

    def __init__(self, s):
        self.s = s
        self.i = 0

    def ws(self):
        while self.i < len(self.s):
            if self.s[self.i].isspace():
                self.i += 1
            elif self.s.startswith("//", self.i):
                while self.i < len(self.s) and self.s[self.i] != "\n":
                    self.i += 1
            else:
                break

    def peek(self, t):
        self.ws()
        return self.s.startswith(t, self.i)

    def eat(self, t):
        self.ws()
        assert self.s.startswith(t, self.i), (t, self.s[self.i:self.i + 60])
        self.i += len(t)

    def seq(self, close):
        out = []
        while True:
            if self.peek(close):
                self.eat(close)
                return out
            out.append(self.value())
            if self.peek(","):
                self.eat(",")

    def suffix(self, v):
        while True:
            self.ws()
            m = re.compile(r"\.(to_vec|to_string|into_boxed_slice|to_owned)\(\)").match(self.s, self.i)
            if m:
                self.i = m.end()
                continue
            return v

    def string(self):
        # after the opening quote
        out = []
        while self.s[self.i] != '"':
            c = self.s[self.i]
            if c == "\\":
                self.i += 1
                c = {"n": "\n", "t": "\t", "\\": "\\", '"': '"', "'": "'"}[self.s[self.i]]
            out.append(c)
            self.i += 1
        self.i += 1
        return "".join(out)

    def value(self):
        self.ws()
        s = self.s
        if s.startswith('b"', self.i):
            self.i += 2
            return self.suffix(self.string())
        if s.startswith('"', self.i):
            self.i += 1
            return self.suffix(self.string())
        if s.startswith("b'", self.i):
            c = s[self.i + 2]
            self.i += 4
            return ord(c)
        if s.startswith("'", self.i):
            c = s[self.i + 1]
            self.i += 3
            return c
        if s.startswith("vec![", self.i):
            self.i += 5
            first = self.seq_or_repeat()
            return self.suffix(first)
        if s.startswith("Box::new(", self.i):
            self.i += 9
            v = self.value()
            self.eat(")")
            return v
        if s.startswith("[", self.i):
            self.i += 1
            return self.suffix(self.seq("]"))
        if s.startswith("(", self.i):
            self.i += 1
            return self.suffix(self.seq(")"))
        if s.startswith("Some(", self.i):
            self.i += 5
            v = self.value()
            self.eat(")")
            return v
        m = re.compile(r"-?\d[\d_]*(\.\d+)?([eE][-+]?\d+)?").match(s, self.i)
        if m:
            self.i = m.end()
            t = m.group(0).replace("_", "")
            m2 = re.compile(r"(u8|u64|usize|f64|i32)").match(s, self.i)
            if m2:
                self.i = m2.end()
            if "." in t or "e" in t or "E" in t:
                return {"f64": t}  # kept as text: compared after float() on both sides
            return int(t)
        m = re.compile(r"[A-Za-z_][A-Za-z_0-9:]*").match(s, self.i)
        assert m, s[self.i:self.i + 80]
        name = m.group(0)
        self.i = m.end()
        if name == "None":
            return None
        if name in ("true", "false"):
            return name == "true"
        if self.peek("{"):
            self.eat("{")
            d = {}
            while not self.peek("}"):
                self.ws()
                fm = re.compile(r"[a-z_0-9]+").match(s, self.i)
                self.i = fm.end()
                self.eat(":")
                d[fm.group(0)] = self.value()
                if self.peek(","):
                    self.eat(",")
            self.eat("}")
            return d
        return name.split("::")[-1]

    def seq_or_repeat(self):
        # vec![x; n] or vec![a, b, ...]
        if self.peek("]"):
            self.eat("]")
            return []
        first = self.value()
        if self.peek(";"):
            self.eat(";")
            n = self.value()
            self.eat("]")
            return [first] * n
        out = [first]
        if self.peek(","):
            self.eat(",")
        out += self.seq("]")
        return out


def src(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def fn_body(text, name):
    """Text of `fn name(...) { ... }` (brace matched, string aware enough for these files)."""
    m = re.search(r"fn %s\s*\(" % re.escape(name), text)
    assert m, name
    i = text.index("{", m.end())
    depth, j = 0, i
    in_str = False
    while True:
        c = text[j]
        if in_str:
            if c == "\\":
                j += 1
            elif c == '"':
                in_str = False
        else:
            if c == '"':
                in_str = True
            elif c == "'" and text[j + 2] == "'":
                j += 2
            elif c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0:
                    return text[i:j + 1]
        j += 1


def values_after(body, marker, count=1, start=0):
    """Parse `count` literal values, each located right after an occurrence of `marker`."""
    out = []
    pos = start
    for _ in range(count):
        pos = body.index(marker, pos) + len(marker)
        p = P(body)
        p.i = pos
        out.append(p.value())
        pos = p.i
    return out, pos


def static_bytes(text, name):
    m = re.search(r"static %s: [^=]*= b\"" % name, text)
    p = P(text)
    p.i = m.end()
    return p.string()


def assert_eq_pairs(body):
    """All `assert_eq!(a, b)` argument pairs in a function body, parsed where both are literals."""
    out = []
    pos = 0
    while True:
        k = body.find("assert_eq!(", pos)
        if k < 0:
            return out
        p = P(body)
        p.i = k + len("assert_eq!(")
        try:
            a = p.value()
            p.eat(",")
            b = p.value()
            out.append((a, b))
        except AssertionError:
            pass
        pos = k + 5


def main():
    G = {}
    # ---- src/lib.rs ----------------------------------------------------------------
    t = src("src/lib.rs")
    b = fn_body(t, "random_base_")
    G["lib.random_base"] = values_after(b, "truth: Vec<u8> = ")[0][0]
    b = fn_body(t, "random_base_diff_") if "fn random_base_diff_" in t else fn_body(t, "random_base_diff")
    # the test re-assigns `truth` four times for A, C, T, G
    b = t[t.index("#[cfg(test)]"):]
    bd = fn_body(b, "random_base_diff_")
    vals, pos = values_after(bd, "truth: Vec<u8> = ")
    more, _ = values_after(bd, "truth = ", 3, pos)
    G["lib.random_base_diff"] = {"nucs": "ACTG", "truth": vals + more}
    G["lib.random_seq"] = values_after(fn_body(b, "random_seq_"), "assert_eq!(")[0][0]

    # ---- src/alignment.rs ----------------------------------------------------------
    t = src("src/alignment.rs")
    tb = t[t.index("#[cfg(test)]"):]
    ed = []
    for m in re.finditer(r'edit_distance\(b"(\w*)", b"(\w*)"\), (\d+)', fn_body(tb, "compute_edit")):
        ed.append([m.group(1), m.group(2), int(m.group(3))])
    G["alignment.edit_distance"] = ed
    al = []
    for m in re.finditer(r'align\(b"(\w*)", b"(\w*)"\),\s*\((\d+), b"([=XID]*)"', fn_body(tb, "align_edit")):
        al.append([m.group(1), m.group(2), int(m.group(3)), m.group(4)])
    G["alignment.align"] = al

    # ---- src/model/error.rs ----------------------------------------------------------
    t = src("src/model/error.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["model.error.random_error"] = re.findall(r'assert_eq!\(b"(\w*)", &random_error\(b"(\w*)"', fn_body(tb, "random_error_"))
    G["model.error.MODEL"] = static_bytes(tb, "MODEL")
    G["model.error.read_error_model"] = [[m.group(1), m.group(2), int(m.group(3))] for m in re.finditer(
        r'add_errors_to_kmer\(b"(\w*)", &mut rng\),\s*\(b"(\w*)"\.to_vec\(\), (\d+)\)', fn_body(tb, "read_error_model"))]
    G["model.error.random"] = {"kmer": "ACAGTTG", "truth": values_after(fn_body(tb, "random"), "assert_eq!(")[0][0]}

    # ---- src/model/quality.rs --------------------------------------------------------
    t = src("src/model/quality.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["model.quality.MODEL"] = static_bytes(tb, "MODEL")
    G["model.quality.read_quality_model"] = [p[1] for p in assert_eq_pairs(fn_body(tb, "read_quality_model"))]
    G["model.quality.random"] = [p[1] for p in assert_eq_pairs(fn_body(tb, "random"))]
    G["model.quality.ideal"] = [p[1] for p in assert_eq_pairs(fn_body(tb, "ideal"))] if "fn ideal" in tb else None
    G["model.quality.src_random"] = fn_body(tb, "random")[:1200]
    if "fn ideal" in tb:
        G["model.quality.src_ideal"] = fn_body(tb, "ideal")[:1500]

    # ---- src/model/glitch.rs ---------------------------------------------------------
    t = src("src/model/glitch.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["model.glitch.get_value"] = {"params": [10000.0, 25.0, 25.0], "truth": values_after(fn_body(tb, "get_value"), "assert_eq!(")[0][0]}

    # ---- src/model/adapter.rs ----------------------------------------------------------
    t = src("src/model/adapter.rs")
    tb = t[t.index("#[cfg(test)]"):]
    vals, _ = values_after(fn_body(tb, "generate"), "assert_eq!(", 2)
    G["model.adapter.generate"] = {"start": "CGATCACAAA", "end": "CAACATATGT", "params": [90.0, 60.0, 50.0, 20.0],
                                   "starts": vals[0], "ends": vals[1]}

    # ---- src/model/length.rs -----------------------------------------------------------
    t = src("src/model/length.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["model.length.get_value"] = {"params": [50.0, 20.0], "truth": values_after(fn_body(tb, "get_value"), "data,")[0][0]}
    G["model.length.shape"] = [float(x) for x in re.findall(r"assert_eq!\(\w+, ([\d.]+)\);", fn_body(tb, "shape"))]

    # ---- src/model/identity.rs ----------------------------------------------------------
    t = src("src/model/identity.rs")
    tb = t[t.index("#[cfg(test)]"):]
    gv = fn_body(tb, "get_value")
    G["model.identity.get_value_src"] = gv[:400]
    G["model.identity.get_value"] = values_after(gv, "data,")[0][0]
    G["model.identity.shape"] = [float(x) for x in re.findall(r"assert_eq!\(\w+, ([\d.]+)\);", fn_body(tb, "shape"))]

    # ---- src/references.rs ------------------------------------------------------------------
    t = src("src/references.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["references.FASTA"] = static_bytes(tb, "FASTA")
    G["references.read_reference"] = values_after(fn_body(tb, "read_reference"), "refs.sequences,")[0][0]
    G["references.get_reference"] = values_after(fn_body(tb, "get_reference"), "assert_eq!(")[0][0]
    G["references.read_reference_adjust"] = values_after(fn_body(tb, "read_reference_adjust"), "assert_eq!(")[0][0]

    # ---- src/simulate/description.rs -----------------------------------------------------------
    t = src("src/simulate/description.rs")
    G["description.tests_src"] = t[t.index("#[cfg(test)]"):]

    # ---- src/simulate/quality.rs ------------------------------------------------------------------
    t = src("src/simulate/quality.rs")
    tb = t[t.index("#[cfg(test)]"):]
    qb = fn_body(tb, "quality")
    G["simulate.quality"] = {
        "MODEL": static_bytes(tb, "MODEL"),
        "raw": values_after(qb, "let raw = ")[0][0],
        "err": values_after(qb, "let err = ")[0][0],
        "qual": values_after(qb, "assert_eq!(b", 1)[0][0] if False else re.search(r'assert_eq!\(b"(.*?)"\.to_vec\(\), qual\)', qb, re.S).group(1),
    }

    # ---- src/simulate/error.rs ----------------------------------------------------------------------
    t = src("src/simulate/error.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["simulate.error.identity2edit"] = values_after(fn_body(tb, "identity2edit"), "edits,")[0][0]
    G["simulate.error.glitches"] = {"params": [1000.0, 25.0, 25.0], "raw_len": 10000,
                                    "truth": values_after(fn_body(tb, "glitches"), "assert_eq!(")[0][0]}
    G["simulate.error.errors"] = {"k": 7, "target": 20.0, "raw_len": 100,
                                  "truth": values_after(fn_body(tb, "errors"), "assert_eq!(")[0][0]}
    ab = fn_body(tb, "all")
    pairs = re.findall(r'assert_eq!\(b"([^"]*)"\.to_vec\(\), (\w+)\);', ab)
    ids = re.findall(r"assert_eq!\(([\d.]+), edit\);", ab)
    G["simulate.error.all"] = {"glitch": [40.0, 5.0, 5.0], "raw_len": 150, "identities": [0.9, 0.85],
                               "reads": [{"err": pairs[0][0], "cigar": pairs[1][0], "real_id": float(ids[0])},
                                         {"err": pairs[2][0], "cigar": pairs[3][0], "real_id": float(ids[1])}]}
    # Change unit tests are kept as source text + the expectations parsed pairwise
    for name in ("build_change", "contain", "overlap", "merge_change", "merge_containment", "changes"):
        G["simulate.error.src." + name] = fn_body(tb, name)

    # ---- src/simulate/fragments.rs -----------------------------------------------------------------------
    t = src("src/simulate/fragments.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["fragments.FASTA"] = static_bytes(tb, "FASTA")
    G["fragments.read_type"] = values_after(fn_body(tb, "read_type"), "assert_eq!(")[0][0]
    G["fragments.chimera"] = values_after(fn_body(tb, "chimera"), "assert_eq!(")[0][0]
    G["fragments.fragments"] = values_after(fn_body(tb, "fragments"), "assert_eq!(")[0][0]
    G["fragments.iter"] = values_after(fn_body(tb, "iter"), "assert_eq!(")[0][0]
    sb = fn_body(tb, "shape")
    G["fragments.shape"] = {
        "scalars": [float(x) for x in re.findall(r"assert_eq!\(\w+, ([\d.]+)\);", sb)],
        "type_count": values_after(sb, "assert_eq!(type_count, ")[0][0],
        "read_per_ref": values_after(sb, "read_per_ref,")[0][0],
        "mean_len_per_ref": [float(x) for x in re.findall(r"lengths\[\d\]\.len\(\) as f64,\s*([\d.]+)", sb)],
        "contigs": [int(x.replace("_", "")) for x in re.findall(r'">([\d_]+)\\n', sb)],
    }

    # ---- src/simulate/mod.rs ---------------------------------------------------------------------------------
    t = src("src/simulate/mod.rs")
    tb = t[t.index("#[cfg(test)]"):]
    G["simulate.FASTA"] = static_bytes(tb, "FASTA")
    G["simulate.junk_seq"] = re.search(r'assert_eq!\(b"(\w+)"', fn_body(tb, "junk_seq")).group(1)
    G["simulate.random_seq"] = re.search(r'assert_eq!\(\s*b"(\w+)"', fn_body(tb, "random_seq")).group(1)
    G["simulate.read_reference"] = values_after(fn_body(tb, "read_reference"), "assert_eq!(")[0][0]
    G["simulate.produce_read"] = values_after(fn_body(tb, "produce_read"), "assert_eq!(")[0][0]
    G["simulate.shape"] = [float(x) for x in re.findall(r"assert_eq!\(\w+, ([\d.]+)\);", fn_body(tb, "shape"))]

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes;", len(G), "entries")


if __name__ == "__main__":
    sys.exit(main())
