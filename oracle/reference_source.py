"""TEST INFRASTRUCTURE ONLY -- reads the reference's Rust SOURCES and meshes chunks from what it parsed.

The reference (DespawnEngine v0.1.0) is Rust; this image has no rustc, and the reference holds no tests or
golden vectors for the chunk -> mesh path (SURVEY.md sections 4, 8c).  The mechanical pin that is left is the
source text itself.  This module parses, at run time, from ``/root/reference/src``:

  engine/rendering/cube.rs             the six ``pub const *_FACE: [BlockVertex; 6]`` tables (positions, tex_coords)
  content/world/chunks/chunk_mesh.rs   coordinate derivation, the six ``let *_air = ...;`` conditions, the order of
                                       the ``if *_air { vertices.extend(cube::*_FACE ...`` blocks, the ``map_uv``
                                       closure, ``pos_offset``, the fallback AtlasUV, the early-outs
  content/world/chunks/chunk.rs        CHUNK_SIZE, MAX_CHUNK_INDEX, ``index()``, ``is_air_at_idx``, the generators
  content/world/world.rs               terrain choice by chunk Y in ``get_chunk``, ``to_chunk_coord``
  utils/math.rs                        Vec3 From<[i32;3]> / Add / AddAssign / Mul<f32> are component-wise f32

and evaluates the parsed Rust expressions (translated token by token to numpy, float32 throughout, one rounding
per operation) to produce vertex streams.  It shares NO table and NO code with ``oracle/mesher_oracle.cpp``,
``oracle/numpy_oracle.py`` or the CUDA kernels: face templates, emit conditions, emission order, index formula and
terrain rule all come out of the parsed text.  ``tests/test_oracle_vs_reference_source.py`` holds the oracle to it
byte for byte, and ``tests/golden/make_golden_from_reference.py`` writes the committed fixtures from it (the GPU
box has no ``/root/reference``; there the fixtures stand in for it).

Nothing here is copied from the reference: the Rust text is read from where it lies and interpreted.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

DEFAULT_ROOT = "/root/reference"
FILES = {
    "cube": "src/engine/rendering/cube.rs",
    "chunk_mesh": "src/content/world/chunks/chunk_mesh.rs",
    "chunk": "src/content/world/chunks/chunk.rs",
    "world": "src/content/world/world.rs",
    "math": "src/utils/math.rs",
    "vertex": "src/engine/rendering/vertex.rs",
}

VERTEX_DTYPE = np.dtype([("position", np.float32, 3), ("tex_coords", np.float32, 2)])


def available(root: str = DEFAULT_ROOT) -> bool:
    return all(os.path.exists(os.path.join(root, p)) for p in FILES.values())


class ParseError(RuntimeError):
    pass


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def _need(m, what: str):
    if not m:
        raise ParseError(f"reference source: could not find {what}")
    return m


def _balanced(src: str, open_at: int, open_ch: str = "{", close_ch: str = "}") -> str:
    """Text between the bracket at src[open_at] and its partner."""
    assert src[open_at] == open_ch
    depth = 0
    for i in range(open_at, len(src)):
        if src[i] == open_ch:
            depth += 1
        elif src[i] == close_ch:
            depth -= 1
            if depth == 0:
                return src[open_at + 1:i]
    raise ParseError("unbalanced brackets in reference source")


def _fn_body(src: str, name: str) -> str:
    m = _need(re.search(r"\bfn\s+" + re.escape(name) + r"\b[^(]*\(", src), f"fn {name}")
    args = _balanced(src, m.end() - 1, "(", ")")
    brace = src.index("{", m.end() + len(args) + 1)  # past the argument list ([i32; 3] holds a `;`) and the return type
    return _balanced(src, brace)


# ---- Rust expression -> numpy --------------------------------------------------------------------------------------
_TOKEN = re.compile(r"\s*(\|\||&&|==|!=|<=|>=|[-+*/%()<>\[\],.]|[A-Za-z_][A-Za-z_0-9]*|\d+\.\d*|\d+)")


def _tokens(expr: str) -> List[str]:
    out, i = [], 0
    expr = expr.strip()
    while i < len(expr):
        m = _TOKEN.match(expr, i)
        if not m:
            raise ParseError(f"cannot tokenise Rust expression at {expr[i:i + 20]!r}")
        out.append(m.group(1))
        i = m.end()
    return out


class _Expr:
    """Recursive-descent evaluator for the small expression subset the hot path uses, with Rust's precedence:
    ``||`` < ``&&`` < comparisons < ``+ -`` < ``* / %`` < ``as`` casts < postfix (call, index, field).
    Values are numpy arrays or scalars; integer division is floor division on non-negative operands (usize)."""

    def __init__(self, toks: List[str], env: Dict[str, object]):
        self.t, self.i, self.env = toks, 0, env

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self, want=None):
        tok = self.peek()
        if tok is None or (want is not None and tok != want):
            raise ParseError(f"expected {want!r}, found {tok!r} in Rust expression")
        self.i += 1
        return tok

    def parse(self):
        v = self.or_()
        if self.peek() is not None:
            raise ParseError(f"trailing tokens in Rust expression: {self.t[self.i:]}")
        return v

    def or_(self):
        v = self.and_()
        while self.peek() == "||":
            self.take()
            v = np.logical_or(v, self.and_())  # the right side is evaluated on clamped indices, see is_air below
        return v

    def and_(self):
        v = self.cmp()
        while self.peek() == "&&":
            self.take()
            v = np.logical_and(v, self.cmp())
        return v

    def cmp(self):
        v = self.add()
        ops = {"==": np.equal, "!=": np.not_equal, "<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal}
        if self.peek() in ops:
            op = ops[self.take()]
            v = op(v, self.add())
        return v

    def add(self):
        v = self.mul()
        while self.peek() in ("+", "-"):
            if self.take() == "+":
                v = v + self.mul()
            else:
                v = v - self.mul()
        return v

    def mul(self):
        v = self.cast()
        while self.peek() in ("*", "/", "%"):
            op = self.take()
            r = self.cast()
            if op == "*":
                v = v * r
            elif op == "/":
                v = v / r if _is_float(v) or _is_float(r) else v // r
            else:
                v = v % r
        return v

    def cast(self):
        v = self.postfix()
        while self.peek() == "as":
            self.take()
            ty = self.take()
            if ty == "f32":
                v = np.asarray(v).astype(np.float32) if isinstance(v, np.ndarray) else np.float32(v)
            elif ty in ("usize", "i32", "i64", "u16", "u32"):
                v = np.asarray(v).astype(np.int64) if isinstance(v, np.ndarray) else int(v)
            else:
                raise ParseError(f"unsupported cast `as {ty}`")
        return v

    def postfix(self):
        v = self.atom()
        while True:
            tok = self.peek()
            if tok == "[":
                self.take()
                k = self.or_()
                self.take("]")
                v = v[k] if isinstance(k, np.ndarray) else v[int(k)]  # `blocks[idx]` with an index array, else a literal subscript
            elif tok == ".":
                self.take()
                name = self.take()
                v = v[name] if isinstance(v, dict) else getattr(v, name)
            elif tok == "(" and callable(v):
                self.take()
                args = []
                if self.peek() != ")":
                    args.append(self.or_())
                    while self.peek() == ",":
                        self.take()
                        args.append(self.or_())
                self.take(")")
                v = v(*args)
            else:
                return v

    def atom(self):
        tok = self.take()
        if tok == "(":
            v = self.or_()
            self.take(")")
            return v
        if tok == "[":  # array literal
            items = []
            if self.peek() != "]":
                items.append(self.or_())
                while self.peek() == ",":
                    self.take()
                    if self.peek() == "]":
                        break
                    items.append(self.or_())
            self.take("]")
            return items
        if tok == "-":
            return -self.atom()
        if re.fullmatch(r"\d+\.\d*", tok):
            return np.float32(tok)
        if tok.isdigit():
            return int(tok)
        if tok in self.env:
            return self.env[tok]
        raise ParseError(f"unknown identifier {tok!r} in Rust expression")


def _is_float(v) -> bool:
    return isinstance(v, (float, np.floating)) or (isinstance(v, np.ndarray) and v.dtype.kind == "f")


def rust_eval(expr: str, env: Dict[str, object]):
    return _Expr(_tokens(expr), env).parse()


# ---- the parsed reference ------------------------------------------------------------------------------------------
@dataclass
class ReferenceSource:
    root: str
    chunk_size: int
    max_chunk_index: int
    air_block_id: str
    index_expr: str                      # Chunk::index body
    is_air_expr: str                     # Chunk::is_air_at_idx body
    coord_exprs: Dict[str, str]          # x, y, z from idx (chunk_mesh.rs)
    air_conditions: Dict[str, str]       # "top" -> Rust expression, in source order
    emission: List[Tuple[str, str]]      # [(condition name, FACE const name)] in source order
    templates: Dict[str, np.ndarray]     # FACE const name -> (6, 5) float32: px, py, pz, u, v
    map_uv_exprs: List[str]              # two Rust expressions (u, v)
    fallback_uv: Tuple[List[float], List[float]]
    pos_offset_expr: str
    early_out_palette_len: int           # `if chunk.palette.len() == N { return None; }`
    loop_range: Tuple[str, str]          # `for idx in A..B`
    terrain_rule: List[Tuple[str, str, str]]  # [(condition or "else", generator fn, block id or "")]
    generators: Dict[str, str]           # generator fn -> Rust expression of the block id chosen per voxel
    vertex_fields: List[Tuple[str, str]] = field(default_factory=list)

    # -- chunk.rs ------------------------------------------------------------------------------------------------
    def index(self, x, y, z):
        return rust_eval(self.index_expr, {"x": x, "y": y, "z": z, "CHUNK_SIZE": self.chunk_size})

    # -- world.rs:28-47 + chunk.rs generators: global ids with "template:dirt" (or whatever the source names) = 1 -------
    def reference_terrain(self, pos, block_ids: Dict[str, int] | None = None) -> np.ndarray:
        """Voxels of World::get_chunk(pos) as GLOBAL ids.  `block_ids` maps the block-id strings the source uses to
        numbers (default: air -> 0, first other string met -> 1, ...)."""
        ids = dict(block_ids) if block_ids else {}
        ids.setdefault(self.air_block_id, 0)
        chosen = None
        for cond, gen, arg in self.terrain_rule:
            if cond == "else" or bool(rust_eval(cond, {"pos": [int(p) for p in pos]})):
                chosen = (gen, arg)
                break
        if chosen is None:
            raise ParseError("terrain rule does not cover this position")
        gen, arg = chosen
        n = self.chunk_size
        z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        out = np.zeros(n * n * n, dtype=np.uint16)
        if arg and arg not in ids:
            ids[arg] = max(ids.values()) + 1
        # the generator's per-voxel choice: a Rust `if c { a } else { b }` or a single name
        choice = self.generators[gen]
        m = re.fullmatch(r"if\s+(.*?)\s*\{\s*(\w+)\s*\}\s*else\s*\{\s*(\w+)\s*\}", choice.strip(), flags=re.S)
        env = {"x": x, "y": y, "z": z, "CHUNK_SIZE": n}
        names = {"dirt_id": ids.get(arg, 0), "AIR_BLOCK_ID": ids[self.air_block_id]}
        if m:
            cond_v = rust_eval(m.group(1), env)
            vals = np.where(cond_v, names[m.group(2)], names[m.group(3)])
        else:
            vals = np.full(x.shape, names[choice.strip()])
        out[np.asarray(self.index(x, y, z)).reshape(-1)] = vals.reshape(-1).astype(np.uint16)
        # set_block stores 0 for AIR_BLOCK_ID and the palette index otherwise (chunk.rs:49-68); with one non-air
        # block per generator the palette index is 1 == the global id used here
        return out

    # -- chunk_mesh.rs:13-129 ------------------------------------------------------------------------------------------
    def mesh_chunk(self, pos, ids: np.ndarray, uv_table: np.ndarray) -> np.ndarray:
        """The vertex stream build_chunk_mesh produces for a chunk at `pos` whose voxels (global ids, layout given
        by the parsed index()) are `ids`; `uv_table[id] = (uv_min.x, uv_min.y, uv_max.x, uv_max.y)`, ids beyond the
        table use the fallback parsed from the source.  Empty result <=> the reference returns None."""
        n = self.chunk_size
        ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(-1)
        assert ids.shape[0] == n ** 3
        uv_table = np.ascontiguousarray(uv_table, dtype=np.float32).reshape(-1, 4)
        # `if chunk.palette.len() == 1 { return None }`: the palette holds air plus every distinct block id set
        if 1 + len(set(ids[ids != 0].tolist())) == self.early_out_palette_len:
            return np.empty(0, dtype=VERTEX_DTYPE)
        lo = int(rust_eval(self.loop_range[0], {"MAX_CHUNK_INDEX": self.max_chunk_index}))
        hi = int(rust_eval(self.loop_range[1], {"MAX_CHUNK_INDEX": self.max_chunk_index}))
        idx = np.arange(lo, hi, dtype=np.int64)

        def is_air_at(i):
            # The Rust `a || is_air(i)` never evaluates is_air when `a` holds (usize would underflow / the Vec would
            # be out of bounds).  Vectorised, the right side IS evaluated everywhere, so clamp the index: wherever
            # the clamp changes it, the left side is true and the OR discards this value.
            i = np.clip(np.asarray(i, dtype=np.int64), 0, ids.shape[0] - 1)
            return rust_eval(self.is_air_expr, {"self": {"blocks": ids}, "idx": i})

        env = {"idx": idx, "CHUNK_SIZE": n, "is_air": is_air_at}
        for name, e in self.coord_exprs.items():
            env[name] = rust_eval(e, env)
        solid = ~np.asarray(is_air_at(idx), dtype=bool)
        flags = np.stack([np.asarray(rust_eval(self.air_conditions[c], env), dtype=bool) & solid for c, _ in self.emission], axis=1)
        vox, which = np.nonzero(flags)  # row-major: voxel ascending, then the source order of the `if` blocks
        if vox.size == 0:
            return np.empty(0, dtype=VERTEX_DTYPE)
        # pos_offset (f32): evaluated from the parsed expression with Vec3 arithmetic component-wise
        px = {k: np.asarray(env[k])[vox] for k in ("x", "y", "z")}
        off = rust_eval(self.pos_offset_expr, {
            "Vec3": _Vec3Namespace(), "x": px["x"], "y": px["y"], "z": px["z"],
            "chunk": {"position": [np.int32(p) for p in pos]}, "CHUNK_SIZE": n})
        tmpl = np.stack([self.templates[f] for _, f in self.emission])[which]  # (F, 6, 5)
        out = np.empty((vox.size, 6), dtype=VERTEX_DTYPE)
        for k in range(3):  # v.position += pos_offset  (math.rs AddAssign: component-wise f32 add)
            out["position"][:, :, k] = tmpl[:, :, k] + off.c[k][:, None].astype(np.float32)
        bid = ids[idx[vox]].astype(np.int64)
        known = bid < uv_table.shape[0]
        rows = np.where(known[:, None], uv_table[np.minimum(bid, uv_table.shape[0] - 1)],
                        np.array(self.fallback_uv[0] + self.fallback_uv[1], dtype=np.float32)[None, :]).astype(np.float32)
        for k in range(2):  # v.tex_coords = map_uv(v.tex_coords, atlas)
            atlas = {"uv_min": [rows[:, 0:1], rows[:, 1:2]], "uv_max": [rows[:, 2:3], rows[:, 3:4]]}
            orig = [tmpl[:, :, 3], tmpl[:, :, 4]]
            val = rust_eval(self.map_uv_exprs[k], {"atlas": atlas, "orig": orig})
            assert val.dtype == np.float32
            out["tex_coords"][:, :, k] = val
        return out.reshape(-1)


class _V3:
    def __init__(self, c):
        self.c = [np.asarray(v, dtype=np.float32) for v in c]

    def __add__(self, o):  # math.rs `impl Add for Vec3`
        return _V3([a + b for a, b in zip(self.c, o.c)])

    def __mul__(self, s):  # math.rs `impl Mul<f32> for Vec3`
        assert _is_float(s)
        return _V3([a * np.float32(s) for a in self.c])


class _Vec3Namespace:
    @staticmethod
    def from_(v):
        # From<[f32;3]> is the identity, From<[i32;3]> casts every component `as f32`
        return _V3([np.asarray(c).astype(np.float32) for c in v])


def _parse_templates(cube_src: str) -> Dict[str, np.ndarray]:
    out = {}
    for m in re.finditer(r"pub\s+const\s+(\w+_FACE)\s*:\s*\[\s*BlockVertex\s*;\s*(\d+)\s*\]\s*=\s*\[", cube_src):
        body = _balanced(cube_src, m.end() - 1, "[", "]")
        verts = re.findall(r"BlockVertex\s*\{\s*position\s*:\s*Vec3\(\[\s*([^\]]+?)\s*\]\)\s*,\s*tex_coords\s*:\s*\[\s*([^\]]+?)\s*\]\s*,?\s*\}", body)
        if len(verts) != int(m.group(2)):
            raise ParseError(f"{m.group(1)}: expected {m.group(2)} vertices, parsed {len(verts)}")
        rows = [[np.float32(t) for t in p.split(",")] + [np.float32(t) for t in uv.split(",")] for p, uv in verts]
        out[m.group(1)] = np.array(rows, dtype=np.float32)
    return out


def parse(root: str = DEFAULT_ROOT) -> ReferenceSource:
    src = {k: _strip_comments(open(os.path.join(root, p)).read()) for k, p in FILES.items()}

    # chunk.rs
    chunk_size = int(_need(re.search(r"pub\s+const\s+CHUNK_SIZE\s*:\s*usize\s*=\s*(\d+)\s*;", src["chunk"]), "CHUNK_SIZE").group(1))
    max_idx = int(_need(re.search(r"pub\s+const\s+MAX_CHUNK_INDEX\s*:\s*usize\s*=\s*(\d+)\s*;", src["chunk"]), "MAX_CHUNK_INDEX").group(1))
    air_id = _need(re.search(r'pub\s+const\s+AIR_BLOCK_ID\s*:\s*&str\s*=\s*"([^"]+)"\s*;', src["chunk"]), "AIR_BLOCK_ID").group(1)
    index_expr = _fn_body(src["chunk"], "index").strip()
    is_air_expr = _fn_body(src["chunk"], "is_air_at_idx").strip()
    generators = {}
    for gen in ("generate_flat", "generate_full", "generate_empty"):
        body = _fn_body(src["chunk"], gen)
        loops = re.findall(r"for\s+(\w)\s+in\s+0\s*\.\.\s*CHUNK_SIZE", body)
        if sorted(loops) != ["x", "y", "z"]:
            raise ParseError(f"{gen}: expected loops over x, y, z in 0..CHUNK_SIZE")
        m = re.search(r"let\s+block_id\s*=\s*(if\b.*?\}\s*else\s*\{.*?\})\s*;", body, flags=re.S)
        call = _need(re.search(r"self\.set_block\(\s*x\s*,\s*y\s*,\s*z\s*,\s*(\w+)\s*\)", body), f"{gen}: set_block call").group(1)
        generators[gen] = m.group(1) if (m and call == "block_id") else call
    set_block = _fn_body(src["chunk"], "set_block")
    _need(re.search(r"if\s+block_id\s*==\s*AIR_BLOCK_ID\s*\{\s*self\.blocks\[idx\]\s*=\s*0\s*;", set_block), "set_block: air stores palette index 0")

    # world.rs get_chunk: if pos[1] > 0 {generate_empty} else if pos[1] == 0 {generate_flat("..")} else {generate_full("..")}
    body = _fn_body(src["world"], "get_chunk")
    rule = []
    chain = _need(re.search(r"(if\s+pos\[.*)\n\s*let\s+chunk_arc", body, flags=re.S), "get_chunk: terrain rule").group(1)
    for m in re.finditer(r"(?:(else\s+if|if)\s+(.*?)|(else))\s*\{\s*chunk\.(generate_\w+)\(\s*(?:\"([^\"]*)\"\s*,\s*)?content\s*\)\s*;\s*\}", chain, flags=re.S):
        rule.append(("else" if m.group(3) else m.group(2).strip(), m.group(4), m.group(5) or ""))
    if len(rule) < 2 or rule[-1][0] != "else":
        raise ParseError("get_chunk: could not parse the if / else-if / else terrain chain")
    tcc = _fn_body(src["world"], "to_chunk_coord")
    _need(re.search(r"world_coord\.div_euclid\(CHUNK_SIZE as i32\)", tcc), "to_chunk_coord: div_euclid")
    _need(re.search(r"world_coord\.rem_euclid\(CHUNK_SIZE as i32\)", tcc), "to_chunk_coord: rem_euclid")

    # cube.rs
    templates = _parse_templates(src["cube"])

    # chunk_mesh.rs
    body = _fn_body(src["chunk_mesh"], "build_chunk_mesh")
    early = int(_need(re.search(r"if\s+chunk\.palette\.len\(\)\s*==\s*(\d+)\s*\{\s*return\s+None\s*;", body), "palette early-out").group(1))
    loop = _need(re.search(r"for\s+idx\s+in\s+(.+?)\s*\.\.\s*(.+?)\s*\{", body), "voxel loop")
    _need(re.search(r"if\s+chunk\.is_air_at_idx\(idx\)\s*\{\s*continue\s*;", body), "air voxels are skipped")
    _need(re.search(r"let\s+is_air\s*=\s*\|idx:\s*usize\|\s*chunk\.is_air_at_idx\(idx\)\s*;", body), "is_air closure")
    coords = {k: _need(re.search(r"let\s+" + k + r"\s*=\s*([^;]+);", body), f"let {k}").group(1).strip() for k in ("x", "y", "z")}
    conds = {m.group(1): m.group(2).strip() for m in re.finditer(r"let\s+(\w+)_air\s*=\s*(?![\s|])([^;]+);", body)}  # (not the `is_air` closure)
    emission = [(m.group(1), m.group(2)) for m in re.finditer(
        r"if\s+(\w+)_air\s*\{\s*vertices\.extend\(\s*cube::(\w+_FACE)\.map\(\s*\|mut v\|\s*\{\s*v\.position\s*\+=\s*pos_offset\s*;\s*"
        r"v\.tex_coords\s*=\s*map_uv\(\s*v\.tex_coords\s*,\s*atlas\s*\)\s*;\s*v\s*\}\s*\)\s*\)\s*;\s*\}", body)]
    if len(emission) != len(conds) or {c for c, _ in emission} != set(conds):
        raise ParseError(f"emission blocks {emission} do not match the conditions {list(conds)}")
    for _, f in emission:
        if f not in templates:
            raise ParseError(f"cube::{f} used by build_chunk_mesh but not found in cube.rs")
    m = _need(re.search(r"let\s+map_uv\s*=\s*\|orig:\s*\[f32;\s*2\],\s*atlas:\s*AtlasUV\|\s*->\s*\[f32;\s*2\]\s*\{\s*\[(.*?)\]\s*\}\s*;", body, flags=re.S), "map_uv closure")
    map_uv = [e.strip() for e in _split_top_level(m.group(1))]
    if len(map_uv) != 2:
        raise ParseError("map_uv: expected two components")
    fb = _need(re.search(r"unwrap_or\(\s*AtlasUV\s*\{\s*uv_min\s*:\s*\[([^\]]+)\]\s*,\s*uv_max\s*:\s*\[([^\]]+)\]\s*,?\s*\}\s*\)", body), "fallback AtlasUV")
    fallback = ([float(t) for t in fb.group(1).split(",")], [float(t) for t in fb.group(2).split(",")])
    pos_off = _need(re.search(r"let\s+pos_offset\s*=\s*([^;]+);", body), "pos_offset").group(1)
    pos_off = re.sub(r"Vec3::from\(", "Vec3.from_(", " ".join(pos_off.split()))  # `::` is not in the token set
    _need(re.search(r"if\s+!vertices\.is_empty\(\)", body), "empty vertex list -> None")

    # math.rs: the Vec3 operations the path uses are component-wise single f32 operations
    for impl_, pat in (("From<[i32; 3]>", r"Vec3\(\[\s*value\[0\] as f32\s*,\s*value\[1\] as f32\s*,\s*value\[2\] as f32\s*\]\)"),
                       ("AddAssign", r"self\.0\[0\]\s*\+=\s*rhs\.0\[0\];\s*self\.0\[1\]\s*\+=\s*rhs\.0\[1\];\s*self\.0\[2\]\s*\+=\s*rhs\.0\[2\];"),
                       ("Add", r"self\.0\[0\]\s*\+\s*rhs\.0\[0\]\s*,\s*self\.0\[1\]\s*\+\s*rhs\.0\[1\]\s*,\s*self\.0\[2\]\s*\+\s*rhs\.0\[2\]"),
                       ("Mul<f32>", r"self\.0\[0\]\s*\*\s*rhs\s*,\s*self\.0\[1\]\s*\*\s*rhs\s*,\s*self\.0\[2\]\s*\*\s*rhs")):
        _need(re.search(pat, src["math"]), f"math.rs: component-wise {impl_} for Vec3")

    # vertex.rs: #[repr(C)] { position: Vec3, tex_coords: [f32; 2] }
    vs = _need(re.search(r"#\[repr\(C\)\]\s*pub\s+struct\s+BlockVertex\s*\{(.*?)\}", src["vertex"], flags=re.S), "BlockVertex").group(1)
    fields = re.findall(r"pub\s+(\w+)\s*:\s*([^,]+),", re.sub(r"#\[[^\]]*\]", "", vs))
    if [(a, " ".join(b.split())) for a, b in fields] != [("position", "Vec3"), ("tex_coords", "[f32; 2]")]:
        raise ParseError(f"BlockVertex layout changed: {fields}")

    return ReferenceSource(root=root, chunk_size=chunk_size, max_chunk_index=max_idx, air_block_id=air_id, index_expr=index_expr,
                           is_air_expr=is_air_expr, coord_exprs=coords, air_conditions=conds, emission=emission, templates=templates,
                           map_uv_exprs=map_uv, fallback_uv=fallback, pos_offset_expr=pos_off, early_out_palette_len=early,
                           loop_range=(loop.group(1), loop.group(2)), terrain_rule=rule, generators=generators, vertex_fields=fields)


def _split_top_level(s: str) -> List[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out
