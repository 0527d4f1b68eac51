"""Triangle extraction from Open Inventor 2.x / VRML 1.0 / VRML 2.0 ascii scene graphs.

GraspIt! reads geometry through Coin3D (``SoDB::readAll`` / ``readAllVRML``,
reference ``src/body.cpp:456-481``) and extracts triangles with an ``SoCallbackAction``
whose callback multiplies every vertex by Coin's **fp32** model matrix and drops triangles
with two identical vertices (``src/body.cpp:1225-1249``).  Coin is not available, so this
module re-creates those semantics for the node types GraspIt!'s assets use:

* state nodes: ``Separator``/``TransformSeparator`` (scoped), ``Group``, ``Translation``,
  ``Rotation``, ``RotationXYZ``, ``Scale``, ``Transform``, ``MatrixTransform``,
  ``Coordinate3``, ``VertexProperty``; VRML2 ``Transform``/``Group``/``Shape``/``Coordinate``
* shapes: ``FaceSet``, ``IndexedFaceSet``, ``IndexedTriangleStripSet``, ``TriangleStripSet``,
  ``Cube``/``Box``, ``Cylinder``, ``Cone``, ``Sphere`` (the last three with a fixed tessellation;
  none of the benchmark scenes depend on them)

Matrices follow Coin's ``SbMatrix`` row-vector convention (``p' = p * M``) and every product
is rounded to fp32 step by step, outer node first, like ``SoModelMatrixElement`` does.
The output is an ``(N, 3, 3)`` float32 array -- the exact values GraspIt! widens to double
(SURVEY.md 8(a) row A0).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import List, Optional, Union

import numpy as np

F32 = np.float32

_TOKEN_RE = re.compile(
    r"""\s*(?:
        (?P<comment>\#[^\n]*) |
        (?P<str>"(?:[^"\\]|\\.)*") |
        (?P<punct>[{}\[\],|]) |
        (?P<num>[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)(?![A-Za-z_])) |
        (?P<hex>0x[0-9a-fA-F]+) |
        (?P<id>[^\s{}\[\],|"\#]+)
    )""",
    re.X,
)


def _tokenize(text: str):
    pos = 0
    n = len(text)
    out = []
    while pos < n:
        m = _TOKEN_RE.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise ValueError(f"scene graph: cannot tokenize at offset {pos}: {text[pos:pos+40]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "comment":
            continue
        if kind == "punct":
            if m.group(kind) == ",":
                continue
            out.append(("p", m.group(kind)))
        elif kind == "num":
            out.append(("n", float(m.group(kind))))
        elif kind == "hex":
            out.append(("n", float(int(m.group(kind), 16))))
        elif kind == "str":
            out.append(("s", m.group(kind)[1:-1]))
        else:
            out.append(("i", m.group(kind)))
    return out


@dataclass
class Node:
    type: str
    name: Optional[str] = None
    fields: dict = field(default_factory=dict)      # name -> list of scalars | Node | list[Node]
    children: List["Node"] = field(default_factory=list)


class _Parser:
    def __init__(self, tokens):
        self.t = tokens
        self.i = 0
        self.defs = {}

    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else ("eof", None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def at_node_start(self):
        k, v = self.peek()
        if k != "i":
            return False
        if v in ("DEF", "USE"):
            return True
        return self.peek(1) == ("p", "{")

    def parse_top(self):
        nodes = []
        while self.peek()[0] != "eof":
            if self.at_node_start():
                nodes.append(self.parse_node())
            else:
                self.next()  # stray token (e.g. VRML2 ROUTE statements)
        return nodes

    def parse_node(self) -> Node:
        k, v = self.next()
        name = None
        if v == "USE":
            ref = self.next()[1]
            return self.defs.get(ref, Node("Unknown"))
        if v == "DEF":
            name = self.next()[1]
            k, v = self.next()
        node = Node(type=v, name=name)
        if name is not None:
            self.defs[name] = node
        tok = self.next()
        if tok != ("p", "{"):
            raise ValueError(f"scene graph: expected '{{' after {v}, got {tok}")
        while True:
            k, v = self.peek()
            if k == "eof":
                raise ValueError("scene graph: unexpected end of file")
            if (k, v) == ("p", "}"):
                self.next()
                break
            if self.at_node_start():
                node.children.append(self.parse_node())
                continue
            if k != "i":
                self.next()      # tolerate junk
                continue
            fname = self.next()[1]
            if fname == "fields":   # Inventor extension-node field declarations: fields [ ... ]
                self._skip_bracket()
                continue
            node.fields[fname] = self.parse_value()
        return node

    def _skip_bracket(self):
        if self.peek() == ("p", "["):
            depth = 0
            while True:
                k, v = self.next()
                if (k, v) == ("p", "["):
                    depth += 1
                elif (k, v) == ("p", "]"):
                    depth -= 1
                    if depth == 0:
                        return
                elif k == "eof":
                    return

    def parse_value(self):
        k, v = self.peek()
        if (k, v) == ("p", "["):
            self.next()
            items: list = []
            while True:
                k, v = self.peek()
                if (k, v) == ("p", "]"):
                    self.next()
                    break
                if k == "eof":
                    raise ValueError("scene graph: unterminated list")
                if self.at_node_start():
                    items.append(self.parse_node())
                else:
                    self.next()
                    if k in ("n", "s", "i"):
                        items.append(v)
            return items
        if self.at_node_start():
            return self.parse_node()
        if k == "i":
            # enum / boolean / NULL, possibly OR-ed flags
            vals = [self.next()[1]]
            while self.peek() == ("p", "|"):
                self.next()
                vals.append(self.next()[1])
            return vals
        vals = []
        while self.peek()[0] in ("n", "s"):
            vals.append(self.next()[1])
        return vals


def parse_scene(text: str) -> List[Node]:
    return _Parser(_tokenize(text)).parse_top()


# ----------------------------------------------------------------------------------------
# fp32 SbMatrix-style helpers (row-vector convention)
# ----------------------------------------------------------------------------------------

def _ident():
    return np.eye(4, dtype=F32)


def _mult_left(model: np.ndarray, m: np.ndarray) -> np.ndarray:
    """model <- m * model, all fp32 (SbMatrix::multLeft)."""
    return (m.astype(F32) @ model.astype(F32)).astype(F32)


def _translate(v):
    m = _ident()
    m[3, 0:3] = np.asarray(v, dtype=F32)
    return m


def _scale(v):
    m = _ident()
    m[0, 0], m[1, 1], m[2, 2] = F32(v[0]), F32(v[1]), F32(v[2])
    return m


def _rotation_quat(axis, angle):
    """SbRotation(axis, radians): unit quaternion (x,y,z,w) in fp32."""
    a = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(a)
    if n == 0.0:
        return np.array([0, 0, 0, 1], dtype=F32)
    a = a / n
    s = math.sin(angle / 2.0)
    return np.array([a[0] * s, a[1] * s, a[2] * s, math.cos(angle / 2.0)], dtype=F32)


def _rotate(axis, angle):
    x, y, z, w = [F32(c) for c in _rotation_quat(axis, angle)]
    two = F32(2.0)
    m = _ident()
    m[0, 0] = w * w + x * x - y * y - z * z
    m[0, 1] = two * x * y + two * w * z
    m[0, 2] = two * x * z - two * w * y
    m[1, 0] = two * x * y - two * w * z
    m[1, 1] = w * w - x * x + y * y - z * z
    m[1, 2] = two * y * z + two * w * x
    m[2, 0] = two * x * z + two * w * y
    m[2, 1] = two * y * z - two * w * x
    m[2, 2] = w * w - x * x - y * y + z * z
    return m


def _apply(model: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """SbMatrix::multVecMatrix for an (N,3) fp32 array: dst = src * M (affine, w == 1)."""
    pts = pts.astype(F32)
    m = model.astype(F32)
    out = np.empty_like(pts)
    for c in range(3):
        acc = pts[:, 0] * m[0, c]
        acc = (acc + pts[:, 1] * m[1, c]).astype(F32)
        acc = (acc + pts[:, 2] * m[2, c]).astype(F32)
        out[:, c] = (acc + m[3, c]).astype(F32)
    return out


def _floats(v, n=None, default=None):
    if v is None:
        return default
    vals = [float(x) for x in v if isinstance(x, (int, float))]
    if n is not None and len(vals) < n:
        return default
    return vals


class _State:
    __slots__ = ("model", "coords")

    def __init__(self, model=None, coords=None):
        self.model = _ident() if model is None else model
        self.coords = coords

    def copy(self):
        return _State(self.model.copy(), self.coords)


class TriangleExtractor:
    def __init__(self):
        self.tris: List[np.ndarray] = []     # each (k,3,3) fp32, already transformed

    # -- emit helpers -------------------------------------------------------------------
    def _emit(self, st: _State, verts: np.ndarray, faces):
        if len(faces) == 0:
            return
        f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
        p = _apply(st.model, np.asarray(verts, dtype=F32).reshape(-1, 3))
        self.tris.append(p[f])

    @staticmethod
    def _poly_to_tris(idx):
        # fan triangulation (v0, vi, vi+1): what Coin produces for triangles and quads
        return [(idx[0], idx[i], idx[i + 1]) for i in range(1, len(idx) - 1)]

    # -- traversal ------------------------------------------------------------------------
    def run(self, nodes: List[Node]):
        st = _State()
        for n in nodes:
            self.visit(n, st)
        if not self.tris:
            return np.zeros((0, 3, 3), dtype=F32)
        t = np.concatenate(self.tris, axis=0).astype(F32)
        # drop triangles with two identical (fp32) vertices -- src/body.cpp:1241
        a, b, c = t[:, 0], t[:, 1], t[:, 2]
        degenerate = (a == b).all(1) | (b == c).all(1) | (a == c).all(1)
        return np.ascontiguousarray(t[~degenerate])

    def visit(self, n: Node, st: _State):
        ty = n.type
        f = n.fields
        if ty in ("Separator", "TransformSeparator", "Selection", "Annotation"):
            inner = st.copy()
            for c in n.children:
                self.visit(c, inner)
            if ty == "TransformSeparator":
                st.coords = inner.coords
        elif ty in ("Group", "Switch") and "children" not in f:
            for c in n.children:
                self.visit(c, st)
        elif ty == "Translation":
            st.model = _mult_left(st.model, _translate(_floats(f.get("translation"), 3, [0, 0, 0])))
        elif ty == "Scale":
            st.model = _mult_left(st.model, _scale(_floats(f.get("scaleFactor"), 3, [1, 1, 1])))
        elif ty == "RotationXYZ":
            axis = {"X": (1, 0, 0), "Y": (0, 1, 0), "Z": (0, 0, 1)}[str(f.get("axis", ["X"])[0])]
            ang = _floats(f.get("angle"), 1, [0.0])[0]
            st.model = _mult_left(st.model, _rotate(axis, ang))
        elif ty == "Rotation":
            r = _floats(f.get("rotation"), 4, [0, 0, 1, 0])
            st.model = _mult_left(st.model, _rotate(r[0:3], r[3]))
        elif ty == "MatrixTransform":
            m = np.asarray(_floats(f.get("matrix"), 16, list(np.eye(4).ravel())), dtype=F32).reshape(4, 4)
            st.model = _mult_left(st.model, m)
        elif ty == "Transform" and "children" not in f:
            self._inventor_transform(f, st)
        elif ty == "Coordinate3":
            st.coords = np.asarray(_floats(f.get("point"), None, []), dtype=F32).reshape(-1, 3)
        elif ty == "VertexProperty":
            if "vertex" in f:
                st.coords = np.asarray(_floats(f["vertex"]), dtype=F32).reshape(-1, 3)
        elif ty == "FaceSet":
            self._face_set(f, st)
        elif ty in ("IndexedFaceSet",) and "coord" not in f:
            self._indexed_face_set(f, st)
        elif ty == "IndexedTriangleStripSet":
            self._indexed_strip_set(f, st)
        elif ty == "TriangleStripSet":
            self._strip_set(f, st)
        elif ty in ("Cube",):
            w = _floats(f.get("width"), 1, [2.0])[0]
            h = _floats(f.get("height"), 1, [2.0])[0]
            d = _floats(f.get("depth"), 1, [2.0])[0]
            self._box(st, w, h, d)
        elif ty == "Cylinder":
            self._cylinder(st, _floats(f.get("radius"), 1, [1.0])[0], _floats(f.get("height"), 1, [2.0])[0], cone=False)
        elif ty == "Cone":
            self._cylinder(st, _floats(f.get("bottomRadius"), 1, [1.0])[0], _floats(f.get("height"), 1, [2.0])[0], cone=True)
        elif ty == "Sphere":
            self._sphere(st, _floats(f.get("radius"), 1, [1.0])[0])
        # ---------------- VRML 2.0 ----------------
        elif ty in ("Transform", "Group", "Collision", "Anchor", "Billboard") and "children" in f:
            inner = st.copy()
            if ty == "Transform":
                self._vrml2_transform(f, inner)
            ch = f["children"]
            for c in (ch if isinstance(ch, list) else [ch]):
                if isinstance(c, Node):
                    self.visit(c, inner)
        elif ty == "Shape":
            g = f.get("geometry")
            if isinstance(g, Node):
                self.visit(g, st.copy())
        elif ty == "IndexedFaceSet":   # VRML2 flavour (coord field)
            c = f.get("coord")
            inner = st.copy()
            if isinstance(c, Node):
                inner.coords = np.asarray(_floats(c.fields.get("point"), None, []), dtype=F32).reshape(-1, 3)
            self._indexed_face_set(f, inner)
        elif ty == "Box":
            s = _floats(f.get("size"), 3, [2, 2, 2])
            self._box(st, s[0], s[1], s[2])
        else:
            # property / unknown nodes: Material, ShapeHints, Normal, Info, Texture2, ...
            for c in n.children:
                self.visit(c, st)

    # -- transforms -----------------------------------------------------------------------
    def _inventor_transform(self, f, st):
        t = _floats(f.get("translation"), 3, [0, 0, 0])
        r = _floats(f.get("rotation"), 4, [0, 0, 1, 0])
        s = _floats(f.get("scaleFactor"), 3, [1, 1, 1])
        so = _floats(f.get("scaleOrientation"), 4, [0, 0, 1, 0])
        c = _floats(f.get("center"), 3, [0, 0, 0])
        self._trs(st, t, r, s, so, c)

    def _vrml2_transform(self, f, st):
        t = _floats(f.get("translation"), 3, [0, 0, 0])
        r = _floats(f.get("rotation"), 4, [0, 0, 1, 0])
        s = _floats(f.get("scale"), 3, [1, 1, 1])
        so = _floats(f.get("scaleOrientation"), 4, [0, 0, 1, 0])
        c = _floats(f.get("center"), 3, [0, 0, 0])
        self._trs(st, t, r, s, so, c)

    @staticmethod
    def _trs(st, t, r, s, so, c):
        # SoTransform::doAction order: translate, center, rotate, scaleOrientation, scale, inverse so, -center
        ident_c = all(v == 0 for v in c)
        ident_so = so[3] == 0
        if any(v != 0 for v in t):
            st.model = _mult_left(st.model, _translate(t))
        if not ident_c:
            st.model = _mult_left(st.model, _translate(c))
        if r[3] != 0:
            st.model = _mult_left(st.model, _rotate(r[0:3], r[3]))
        if any(v != 1 for v in s):
            if not ident_so:
                st.model = _mult_left(st.model, _rotate(so[0:3], so[3]))
            st.model = _mult_left(st.model, _scale(s))
            if not ident_so:
                st.model = _mult_left(st.model, _rotate(so[0:3], -so[3]))
        if not ident_c:
            st.model = _mult_left(st.model, _translate([-c[0], -c[1], -c[2]]))

    # -- shapes ---------------------------------------------------------------------------
    def _coords(self, f, st):
        vp = f.get("vertexProperty")
        if isinstance(vp, Node) and "vertex" in vp.fields:
            return np.asarray(_floats(vp.fields["vertex"]), dtype=F32).reshape(-1, 3)
        return st.coords

    def _face_set(self, f, st):
        coords = self._coords(f, st)
        if coords is None:
            return
        start = int(_floats(f.get("startIndex"), 1, [0])[0])
        nv = _floats(f.get("numVertices"), None, [-1])
        faces = []
        k = start
        for cnt in nv:
            cnt = int(cnt)
            if cnt < 0:
                cnt = len(coords) - k
            faces += self._poly_to_tris(list(range(k, k + cnt)))
            k += cnt
        self._emit(st, coords, faces)

    def _indexed_face_set(self, f, st):
        coords = self._coords(f, st)
        if coords is None:
            return
        idx = [int(v) for v in _floats(f.get("coordIndex"), None, [])]
        faces, cur = [], []
        for v in idx + [-1]:
            if v < 0:
                if len(cur) >= 3:
                    faces += self._poly_to_tris(cur)
                cur = []
            else:
                cur.append(v)
        self._emit(st, coords, faces)

    @staticmethod
    def _strip(cur):
        out = []
        for i in range(len(cur) - 2):
            a, b, c = cur[i], cur[i + 1], cur[i + 2]
            out.append((a, b, c) if i % 2 == 0 else (b, a, c))
        return out

    def _indexed_strip_set(self, f, st):
        coords = self._coords(f, st)
        if coords is None:
            return
        idx = [int(v) for v in _floats(f.get("coordIndex"), None, [])]
        faces, cur = [], []
        for v in idx + [-1]:
            if v < 0:
                faces += self._strip(cur)
                cur = []
            else:
                cur.append(v)
        self._emit(st, coords, faces)

    def _strip_set(self, f, st):
        coords = self._coords(f, st)
        if coords is None:
            return
        start = int(_floats(f.get("startIndex"), 1, [0])[0])
        nv = _floats(f.get("numVertices"), None, [-1])
        faces, k = [], start
        for cnt in nv:
            cnt = int(cnt)
            if cnt < 0:
                cnt = len(coords) - k
            faces += self._strip(list(range(k, k + cnt)))
            k += cnt
        self._emit(st, coords, faces)

    def _box(self, st, w, h, d):
        x, y, z = w / 2.0, h / 2.0, d / 2.0
        v = np.array([[-x, -y, -z], [x, -y, -z], [x, y, -z], [-x, y, -z],
                      [-x, -y, z], [x, -y, z], [x, y, z], [-x, y, z]], dtype=F32)
        quads = [(4, 5, 6, 7), (1, 0, 3, 2), (0, 4, 7, 3), (5, 1, 2, 6), (7, 6, 2, 3), (0, 1, 5, 4)]
        faces = []
        for q in quads:
            faces += self._poly_to_tris(list(q))
        self._emit(st, v, faces)

    def _cylinder(self, st, radius, height, cone, sides=24):
        ang = np.linspace(0, 2 * np.pi, sides, endpoint=False)
        bot = np.stack([radius * np.cos(ang), np.full(sides, -height / 2), -radius * np.sin(ang)], 1)
        verts, faces = [list(p) for p in bot], []
        if cone:
            apex = len(verts); verts.append([0, height / 2, 0])
            for i in range(sides):
                faces.append((i, (i + 1) % sides, apex))
        else:
            top = bot.copy(); top[:, 1] = height / 2
            verts += [list(p) for p in top]
            for i in range(sides):
                j = (i + 1) % sides
                faces += [(i, j, sides + j), (i, sides + j, sides + i)]
            ct = len(verts); verts.append([0, height / 2, 0])
            for i in range(sides):
                faces.append((sides + i, sides + (i + 1) % sides, ct))
        cb = len(verts); verts.append([0, -height / 2, 0])
        for i in range(sides):
            faces.append(((i + 1) % sides, i, cb))
        self._emit(st, np.asarray(verts, dtype=F32), faces)

    def _sphere(self, st, radius, stacks=12, slices=24):
        verts, faces = [], []
        for i in range(stacks + 1):
            th = np.pi * i / stacks
            for j in range(slices):
                ph = 2 * np.pi * j / slices
                verts.append([radius * np.sin(th) * np.cos(ph), radius * np.cos(th), radius * np.sin(th) * np.sin(ph)])
        for i in range(stacks):
            for j in range(slices):
                a = i * slices + j; b = i * slices + (j + 1) % slices
                c = a + slices; d = b + slices
                faces += [(a, c, b), (b, c, d)]
        self._emit(st, np.asarray(verts, dtype=F32), faces)


def triangles_from_text(text: str, pre_model: Optional[np.ndarray] = None) -> np.ndarray:
    """Parse a scene-graph file and return its (N,3,3) fp32 triangle soup.

    ``pre_model`` is an optional fp32 4x4 (row-vector convention) applied *above* the file's
    own graph, used for ``<geometryScaling>`` / ``<geometryOffset>`` (src/body.cpp:330-366).
    """
    ex = TriangleExtractor()
    nodes = parse_scene(text)
    st = _State(pre_model.astype(F32).copy() if pre_model is not None else None)
    for n in nodes:
        ex.visit(n, st)
    out = ex.run([])
    return out


def load_triangles(path: str, pre_model: Optional[np.ndarray] = None) -> np.ndarray:
    with open(path, "r", errors="replace") as fh:
        return triangles_from_text(fh.read(), pre_model)
