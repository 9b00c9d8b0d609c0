"""TEST INFRASTRUCTURE -- a small SPIR-V interpreter for compute shaders.

Purpose: EXECUTE the reference's own cull shaders (compiled from /root/reference/src/Renderer/VulkanShaders/*.comp.glsl
with the reference's bundled glslang, oracle/Makefile target `spv`) on the CPU, so that the C++ restatement in
oracle/cull_oracle.cpp is pinned by outputs of the reference's code and not only by our reading of it.  No Vulkan driver,
SPIR-V tool or GPU is needed: the module is parsed from the .spv words and interpreted one invocation at a time.

Scope: exactly the instruction subset those eight shaders use (scalar / vector float and integer arithmetic, structured
control flow, function calls, storage / uniform / push-constant / physical-storage buffers with explicit layouts, one atomic,
one sampled-image fetch, one storage-image write, seven GLSL.std.450 functions).  Anything else raises NotImplementedError.

Round 2: the same machine also executes the reference's nine D3D12 compute shaders (HlslShaders/CS/*.hlsl compiled by the
same bundled glslang with its HLSL front end, `-D -e csMain`, oracle/Makefile target `hlsl_spv`).  That adds OpSwitch (the
front end's early-return structuring), OpUndef, OpVectorTimesMatrix + RowMajor cbuffer matrices, OpImageFetch
(Texture2D.Load), OpImageTexelPointer (RWBuffer<uint> atomics), OpConvertFToU, shifts and UMax/UMin.  Three operations are
undefined in SPIR-V for some operands and get the semantics the D3D shader instructions they were written for have
(Direct3D 11.3 Functional Specification, the instruction set HLSL SM5/6 integer and conversion ops lower to):
  ConvertFToU           `ftou`: round toward zero, NaN -> 0, negative -> 0, >= 2^32 -> 0xFFFFFFFF (saturating)
  ShiftRight/LeftLogical `ushr` / `ishl`: the shift amount is the low 5 bits of the operand
  ImageFetch            `ld`: any coordinate or mip level outside the resource returns 0 in every component

Arithmetic rules: every floating-point instruction is one IEEE-754 binary32 operation, round-to-nearest, no
contraction.  Where SPIR-V leaves the evaluation order open the order of the reference's own math library is used:
  OpMatrixTimesVector   r_i = ((m[0][i]*v0 + m[1][i]*v1) + m[2][i]*v2) + m[3][i]*v3   (BlitzenMathLibrary/blitMLTypes.h:184-193)
  Length                sqrt((x*x + y*y) + z*z)
  Cross                 GLSL definition, each component a*b - c*d
  Log2                  correctly rounded binary32 result of the real log2 (see `log2_mode`)
Invocations run sequentially in ascending gl_GlobalInvocationID order, so atomic appends come out in ascending invocation
order (a real GPU gives the same multiset in arrival order).
"""
import math
import struct

import numpy as np

f32 = np.float32
u32 = np.uint32

# ---- opcode numbers (Khronos SPIR-V specification, unified1) ----------------------------------------------------------
OP = dict(Name=5, MemberName=6, ExtInstImport=11, ExtInst=12, MemoryModel=14, EntryPoint=15, ExecutionMode=16, Capability=17,
          TypeVoid=19, TypeBool=20, TypeInt=21, TypeFloat=22, TypeVector=23, TypeMatrix=24, TypeImage=25, TypeSampler=26,
          TypeSampledImage=27, TypeArray=28, TypeRuntimeArray=29, TypeStruct=30, TypePointer=32, TypeFunction=33,
          TypeForwardPointer=39, ConstantTrue=41, ConstantFalse=42, Constant=43, ConstantComposite=44, ConstantNull=46,
          Function=54, FunctionParameter=55, FunctionEnd=56, FunctionCall=57, Variable=59, Load=61, Store=62, AccessChain=65,
          InBoundsAccessChain=66, Decorate=71, MemberDecorate=72, VectorShuffle=79, CompositeConstruct=80, CompositeExtract=81,
          ImageSampleExplicitLod=88, ImageWrite=99, ConvertFToU=109, ConvertUToF=112, UConvert=113, Bitcast=124, FNegate=127,
          IAdd=128, FAdd=129, ISub=130, FSub=131, IMul=132, FMul=133, FDiv=136, VectorTimesScalar=142, MatrixTimesVector=145,
          LogicalOr=166, LogicalAnd=167, LogicalNot=168, Select=169, IEqual=170, INotEqual=171, UGreaterThan=172,
          UGreaterThanEqual=174, ULessThan=176, ULessThanEqual=178, FOrdEqual=180, FOrdLessThan=184, FOrdGreaterThan=186,
          FOrdLessThanEqual=188, FOrdGreaterThanEqual=190, AtomicIAdd=234, Phi=245, LoopMerge=246, SelectionMerge=247, Label=248,
          Branch=249, BranchConditional=250, Return=253, ReturnValue=254, CopyLogical=400, Source=3, SourceExtension=4,
          String=7, Line=8, ModuleProcessed=330, ConvertUToPtr=120, ConvertPtrToU=117, ShiftRightLogical=194, ShiftLeftLogical=196,
          BitwiseAnd=199, BitwiseOr=197, Undef=1, Switch=251, ImageFetch=95, ImageTexelPointer=60, VectorTimesMatrix=144,
          Image=100)
OPN = {v: k for k, v in OP.items()}
DEC_BUILTIN, DEC_BINDING, DEC_DESCRIPTOR_SET, DEC_OFFSET, DEC_ARRAY_STRIDE, DEC_MATRIX_STRIDE, DEC_ROW_MAJOR = 11, 33, 34, 35, 6, 7, 4
SC_UNIFORM, SC_INPUT, SC_FUNCTION, SC_PRIVATE, SC_PUSH, SC_STORAGE_BUFFER, SC_UNIFORM_CONSTANT, SC_PHYSICAL = 2, 1, 7, 6, 9, 12, 0, 5349
BUILTIN_GLOBAL_INVOCATION_ID, BUILTIN_WORKGROUP_ID, BUILTIN_LOCAL_INVOCATION_ID, BUILTIN_NUM_WORKGROUPS = 28, 26, 27, 24
GLSL_FABS, GLSL_FLOOR, GLSL_LOG2, GLSL_SQRT, GLSL_FMAX, GLSL_FMIN, GLSL_LENGTH, GLSL_CROSS = 4, 8, 30, 31, 40, 37, 66, 68
GLSL_UMIN, GLSL_UMAX = 38, 41


class Type:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


class MemPtr:
    """Pointer into a byte buffer with an explicit layout (StorageBuffer / Uniform / PushConstant / PhysicalStorageBuffer)."""
    __slots__ = ("buf", "off", "tid", "stride", "mstride")

    def __init__(self, buf, off, tid, stride=None, mstride=None):
        self.buf, self.off, self.tid, self.stride, self.mstride = buf, off, tid, stride, mstride


class VarPtr:
    """Pointer into a Function / Private variable: a python list cell plus an index path."""
    __slots__ = ("cell", "path")

    def __init__(self, cell, path=()):
        self.cell, self.path = cell, path


class Sampler2D:
    """Host-side model of a combined image sampler: callable (u, v, lod) -> float (the .x of the fetch)."""

    def __init__(self, fn):
        self.fn = fn


class StorageImage2D:
    def __init__(self, array):
        self.array = array   # numpy float32 [h, w]


class Texture2D:
    """Texture2D<float4> read with .Load(uint3(x, y, mip)) (OpImageFetch): mips = list of float32 [h, w]; only .r is modelled.
    D3D `ld` rule: a mip level or coordinate outside the resource returns 0."""

    def __init__(self, mips):
        self.mips = mips

    def fetch(self, x, y, lod):
        if not (0 <= lod < len(self.mips)):
            return f32(0)
        img = self.mips[lod]
        if 0 <= x < img.shape[1] and 0 <= y < img.shape[0]:
            return f32(img[y, x])
        return f32(0)


class TexelBufferU32:
    """RWBuffer<uint> (a storage texel buffer, format R32ui): numpy uint8 bytes, one u32 per texel."""

    def __init__(self, bytes_u8):
        self.bytes = bytes_u8


class Module:
    def __init__(self, path, log2_mode="rounded"):
        raw = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(raw) // 4), raw)
        assert w[0] == 0x07230203, "not a SPIR-V module"
        self.log2_mode = log2_mode
        self.types, self.consts, self.names, self.member_names = {}, {}, {}, {}
        self.decor, self.member_decor = {}, {}
        self.variables, self.functions = {}, {}
        self.entry, self.local_size, self.ext_glsl = None, (1, 1, 1), None
        i, cur = 5, None
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            a = w[i + 1:i + wc]
            i += wc
            name = OPN.get(op)
            if cur is not None:
                if name == "FunctionEnd":
                    cur = None
                elif name == "FunctionParameter":
                    cur["params"].append(a[1])
                else:
                    if name == "Label":
                        cur["labels"][a[0]] = len(cur["code"])
                    cur["code"].append((op, a))
                continue
            if name in ("Capability", "MemoryModel", "Source", "SourceExtension", "String", "Line", "ModuleProcessed", "TypeForwardPointer"):
                continue
            if name == "ExtInstImport":
                self.ext_glsl = a[0]
            elif name == "EntryPoint":
                self.entry = a[1]
            elif name == "ExecutionMode":
                if a[1] == 17:
                    self.local_size = tuple(a[2:5])
            elif name == "Name":
                self.names[a[0]] = _str(a[1:])
            elif name == "MemberName":
                self.member_names[(a[0], a[1])] = _str(a[2:])
            elif name == "Decorate":
                self.decor.setdefault(a[0], {})[a[1]] = a[2] if len(a) > 2 else True
            elif name == "MemberDecorate":
                self.member_decor.setdefault((a[0], a[1]), {})[a[2]] = a[3] if len(a) > 3 else True
            elif name == "TypeVoid":
                self.types[a[0]] = Type("void")
            elif name == "TypeBool":
                self.types[a[0]] = Type("bool")
            elif name == "TypeInt":
                self.types[a[0]] = Type("int", width=a[1], signed=a[2])
            elif name == "TypeFloat":
                self.types[a[0]] = Type("float", width=a[1])
            elif name == "TypeVector":
                self.types[a[0]] = Type("vector", elem=a[1], count=a[2])
            elif name == "TypeMatrix":
                self.types[a[0]] = Type("matrix", col=a[1], count=a[2])
            elif name == "TypeImage":
                self.types[a[0]] = Type("image", dim=a[2], sampled=a[6])
            elif name == "TypeSampledImage":
                self.types[a[0]] = Type("sampled_image")
            elif name == "TypeArray":
                self.types[a[0]] = Type("array", elem=a[1], length_id=a[2])
            elif name == "TypeRuntimeArray":
                self.types[a[0]] = Type("rtarray", elem=a[1])
            elif name == "TypeStruct":
                self.types[a[0]] = Type("struct", members=list(a[1:]))
            elif name == "TypePointer":
                self.types[a[0]] = Type("pointer", sc=a[1], pointee=a[2])
            elif name == "TypeFunction":
                self.types[a[0]] = Type("function")
            elif name == "ConstantTrue":
                self.consts[a[1]] = True
            elif name == "ConstantFalse":
                self.consts[a[1]] = False
            elif name == "Constant":
                self.consts[a[1]] = self._scalar_from_words(a[0], a[2:])
            elif name == "ConstantComposite":
                self.consts[a[1]] = [self.consts[x] for x in a[2:]]
            elif name == "Variable":
                self.variables[a[1]] = (a[0], a[2])
            elif name == "Undef":
                self.consts[a[1]] = self.default_value(a[0])
            elif name == "Function":
                cur = {"id": a[1], "params": [], "code": [], "labels": {}}
                self.functions[a[1]] = cur
            else:
                raise NotImplementedError("module-level opcode %s (%d)" % (name, op))

    def _scalar_from_words(self, tid, words):
        t = self.types[tid]
        if t.kind == "float":
            assert t.width == 32
            return f32(struct.unpack("<f", struct.pack("<I", words[0]))[0])
        if t.kind == "int":
            v = words[0] | ((words[1] << 32) if len(words) > 1 else 0)
            return int(v) & ((1 << t.width) - 1)
        raise NotImplementedError(t.kind)

    # ---- layout ----------------------------------------------------------------------------------------------------------
    def type_size(self, tid):
        t = self.types[tid]
        if t.kind in ("float", "int"):
            return t.width // 8
        if t.kind == "vector":
            return self.type_size(t.elem) * t.count
        if t.kind == "pointer":
            return 8
        if t.kind == "struct":
            end = 0
            for k, m in enumerate(t.members):
                off = self.member_decor.get((tid, k), {}).get(DEC_OFFSET, 0)
                if self.types[m].kind != "rtarray":
                    end = max(end, off + self.type_size(m))
            return end
        if t.kind == "array":
            return self.decor[tid][DEC_ARRAY_STRIDE] * self.consts[t.length_id]
        if t.kind == "matrix":
            return 16 * t.count
        raise NotImplementedError(t.kind)

    def default_value(self, tid):
        t = self.types[tid]
        if t.kind == "float":
            return f32(0)
        if t.kind == "int":
            return 0
        if t.kind == "bool":
            return False
        if t.kind == "vector":
            return [self.default_value(t.elem) for _ in range(t.count)]
        if t.kind == "matrix":
            return [self.default_value(t.col) for _ in range(t.count)]
        if t.kind == "struct":
            return [self.default_value(x) for x in t.members]
        if t.kind == "array":
            return [self.default_value(t.elem) for _ in range(self.consts[t.length_id])]
        if t.kind == "pointer":
            return None
        raise NotImplementedError(t.kind)

    def variable_names(self):
        """{name: (id, storage class, set, binding)} of the module-level interface variables."""
        out = {}
        for vid, (ptid, sc) in self.variables.items():
            d = self.decor.get(vid, {})
            nm = self.names.get(vid) or self.names.get(self.types[ptid].pointee) or str(vid)
            out[nm] = (vid, sc, d.get(DEC_DESCRIPTOR_SET), d.get(DEC_BINDING))
        return out


def _str(words):
    b = b"".join(struct.pack("<I", x) for x in words)
    return b.split(b"\0")[0].decode()


class Machine:
    """Executes the entry point of a Module.  Resources are bound by the GLSL variable / block name."""

    def __init__(self, module):
        self.m = module
        self.bound = {}        # variable id -> bytearray-like (numpy uint8) | Sampler2D | StorageImage2D
        self.addr_space = []   # [(base, numpy uint8 array)] for physical storage buffer addresses
        self.instr_count = 0

    def bind(self, name, obj):
        names = self.m.variable_names()
        if name not in names:
            raise KeyError("%s is not an interface variable of this shader (has: %s)" % (name, sorted(names)))
        self.bound[names[name][0]] = obj

    def has(self, name):
        return name in self.m.variable_names()

    def register_address(self, base, array):
        self.addr_space.append((int(base), array))

    def _resolve(self, addr):
        for base, arr in self.addr_space:
            if base <= addr < base + max(len(arr), 1) + 1:
                return arr, addr - base
        raise RuntimeError("physical address %#x is not mapped" % addr)

    # ---- memory access -------------------------------------------------------------------------------------------------
    def _load_mem(self, buf, off, tid, mstride=None):
        m = self.m
        t = m.types[tid]
        if t.kind == "float":
            return f32(np.frombuffer(buf, dtype="<f4", count=1, offset=off)[0])
        if t.kind == "int":
            if t.width == 64:
                return int(np.frombuffer(buf, dtype="<u8", count=1, offset=off)[0])
            return int(np.frombuffer(buf, dtype={8: "u1", 16: "<u2", 32: "<u4"}[t.width], count=1, offset=off)[0])
        if t.kind == "vector":
            es = m.type_size(t.elem)
            return [self._load_mem(buf, off + k * es, t.elem) for k in range(t.count)]
        if t.kind == "matrix":
            rowmajor = isinstance(mstride, tuple)
            ms = (mstride[0] if rowmajor else mstride) or 16
            if rowmajor:   # element (row r, column c) lives at r * stride + c * 4: gather each column
                rows = m.types[t.col].count
                return [[self._load_mem(buf, off + r * ms + c * 4, m.types[t.col].elem) for r in range(rows)] for c in range(t.count)]
            return [self._load_mem(buf, off + c * ms, t.col) for c in range(t.count)]
        if t.kind == "struct":
            out = []
            for k, mt in enumerate(t.members):
                d = m.member_decor.get((tid, k), {})
                out.append(self._load_mem(buf, off + d.get(DEC_OFFSET, 0), mt, _mstride(d)))
            return out
        if t.kind == "array":
            st = m.decor[tid][DEC_ARRAY_STRIDE]
            return [self._load_mem(buf, off + k * st, t.elem) for k in range(m.consts[t.length_id])]
        if t.kind == "pointer":   # buffer reference stored in memory: 64-bit address
            addr = int(np.frombuffer(buf, dtype="<u8", count=1, offset=off)[0])
            b2, o2 = self._resolve(addr)
            return MemPtr(b2, o2, t.pointee)
        raise NotImplementedError("load of " + t.kind)

    def _store_mem(self, buf, off, tid, val):
        m = self.m
        t = m.types[tid]
        if t.kind == "float":
            np.frombuffer(buf, dtype="<f4", count=1, offset=off)[0] = val
        elif t.kind == "int":
            np.frombuffer(buf, dtype={8: "u1", 16: "<u2", 32: "<u4", 64: "<u8"}[t.width], count=1, offset=off)[0] = int(val) & ((1 << t.width) - 1)
        elif t.kind == "vector":
            es = m.type_size(t.elem)
            for k in range(t.count):
                self._store_mem(buf, off + k * es, t.elem, val[k])
        elif t.kind == "struct":
            for k, mt in enumerate(t.members):
                self._store_mem(buf, off + m.member_decor.get((tid, k), {}).get(DEC_OFFSET, 0), mt, val[k])
        else:
            raise NotImplementedError("store of " + t.kind)

    def _default(self, tid):
        return self.m.default_value(tid)

    # ---- execution -----------------------------------------------------------------------------------------------------
    def dispatch(self, groups):
        """Runs groups = (gx, gy, gz) workgroups, invocation by invocation in ascending global id (x fastest)."""
        lx, ly, lz = self.m.local_size
        gx, gy, gz = groups
        with np.errstate(all="ignore"):
            for z in range(gz * lz):
                for y in range(gy * ly):
                    for x in range(gx * lx):
                        self.run_invocation((x, y, z))

    def run_invocation(self, gid):
        m = self.m
        lx, ly, lz = m.local_size
        self.gid = gid
        self.globals = {}
        for vid, (ptid, sc) in m.variables.items():
            if sc in (SC_PRIVATE,):
                self.globals[vid] = VarPtr([self._default(m.types[ptid].pointee)])
            elif sc == SC_INPUT:
                b = m.decor.get(vid, {}).get(DEC_BUILTIN)
                if b == BUILTIN_GLOBAL_INVOCATION_ID:
                    v = [gid[0], gid[1], gid[2]]
                elif b == BUILTIN_WORKGROUP_ID:
                    v = [gid[0] // lx, gid[1] // ly, gid[2] // lz]
                elif b == BUILTIN_LOCAL_INVOCATION_ID:
                    v = [gid[0] % lx, gid[1] % ly, gid[2] % lz]
                else:
                    raise NotImplementedError("builtin %s" % b)
                self.globals[vid] = VarPtr([v])
            elif sc in (SC_UNIFORM, SC_STORAGE_BUFFER, SC_PUSH):
                if vid in self.bound:
                    self.globals[vid] = MemPtr(self.bound[vid], 0, m.types[ptid].pointee)
            elif sc == SC_UNIFORM_CONSTANT:
                if vid in self.bound:
                    self.globals[vid] = VarPtr([self.bound[vid]])
        with np.errstate(all="ignore"):
            self._call(m.entry, [])

    def _call(self, fid, args):
        m, fn = self.m, self.m.functions[fid]
        vals = dict(zip(fn["params"], args))
        code, labels = fn["code"], fn["labels"]
        pc, prev_label, cur_label = 0, None, None
        T, C = m.types, m.consts

        def val(i):
            if i in vals:
                return vals[i]
            if i in C:
                return C[i]
            if i in self.globals:
                return self.globals[i]
            raise KeyError("id %d has no value (unbound resource %s?)" % (i, m.names.get(i)))

        while True:
            op, a = code[pc]
            pc += 1
            self.instr_count += 1
            n = OPN.get(op)
            if n == "Label":
                prev_label, cur_label = cur_label, a[0]
            elif n in ("SelectionMerge", "LoopMerge", "Line"):
                pass
            elif n == "Branch":
                pc = labels[a[0]]
            elif n == "BranchConditional":
                pc = labels[a[1]] if val(a[0]) else labels[a[2]]
            elif n == "Return":
                return None
            elif n == "ReturnValue":
                return val(a[0])
            elif n == "Phi":
                for k in range(2, len(a), 2):
                    if a[k + 1] == prev_label:
                        vals[a[1]] = val(a[k])
                        break
                else:
                    raise RuntimeError("phi without matching predecessor")
            elif n == "Variable":
                vals[a[1]] = VarPtr([val(a[3]) if len(a) > 3 else self._default(T[a[0]].pointee)])
            elif n in ("AccessChain", "InBoundsAccessChain"):
                vals[a[1]] = self._access(val(a[2]), [val(x) for x in a[3:]])
            elif n == "Load":
                vals[a[1]] = self._load(val(a[2]))
            elif n == "Store":
                self._store(val(a[0]), val(a[1]))
            elif n == "CopyLogical":
                vals[a[1]] = _deep(val(a[2]))
            elif n == "FunctionCall":
                vals[a[1]] = self._call(a[2], [val(x) for x in a[3:]])
            elif n == "CompositeConstruct":
                out = []
                for x in a[2:]:
                    v = val(x)
                    if isinstance(v, list) and T[a[0]].kind == "vector":
                        out.extend(v)
                    else:
                        out.append(v)
                vals[a[1]] = out
            elif n == "CompositeExtract":
                v = val(a[2])
                for k in a[3:]:
                    v = v[k]
                vals[a[1]] = v
            elif n == "VectorShuffle":
                cat = list(val(a[2])) + list(val(a[3]))
                vals[a[1]] = [cat[k] for k in a[4:]]
            elif n == "FAdd":
                vals[a[1]] = _map2(lambda x, y: f32(x + y), val(a[2]), val(a[3]))
            elif n == "FSub":
                vals[a[1]] = _map2(lambda x, y: f32(x - y), val(a[2]), val(a[3]))
            elif n == "FMul":
                vals[a[1]] = _map2(lambda x, y: f32(x * y), val(a[2]), val(a[3]))
            elif n == "FDiv":
                vals[a[1]] = _map2(lambda x, y: f32(x / y), val(a[2]), val(a[3]))
            elif n == "FNegate":
                vals[a[1]] = _map1(lambda x: f32(-x), val(a[2]))
            elif n == "VectorTimesScalar":
                s = val(a[3])
                vals[a[1]] = [f32(x * s) for x in val(a[2])]
            elif n == "MatrixTimesVector":
                mat, v = val(a[2]), val(a[3])
                rows = len(mat[0])
                out = []
                for r in range(rows):
                    acc = f32(mat[0][r] * v[0])
                    for c in range(1, len(mat)):
                        acc = f32(acc + f32(mat[c][r] * v[c]))
                    out.append(acc)
                vals[a[1]] = out
            elif n == "IAdd":
                vals[a[1]] = _map2(lambda x, y: (x + y) & self._mask(a[0]), val(a[2]), val(a[3]))
            elif n == "ISub":
                vals[a[1]] = _map2(lambda x, y: (x - y) & self._mask(a[0]), val(a[2]), val(a[3]))
            elif n == "IMul":
                vals[a[1]] = _map2(lambda x, y: (x * y) & self._mask(a[0]), val(a[2]), val(a[3]))
            elif n == "IEqual":
                vals[a[1]] = _map2(lambda x, y: x == y, val(a[2]), val(a[3]))
            elif n == "INotEqual":
                vals[a[1]] = _map2(lambda x, y: x != y, val(a[2]), val(a[3]))
            elif n == "ULessThan":
                vals[a[1]] = _map2(lambda x, y: x < y, val(a[2]), val(a[3]))
            elif n == "ULessThanEqual":
                vals[a[1]] = _map2(lambda x, y: x <= y, val(a[2]), val(a[3]))
            elif n == "UGreaterThan":
                vals[a[1]] = _map2(lambda x, y: x > y, val(a[2]), val(a[3]))
            elif n == "UGreaterThanEqual":
                vals[a[1]] = _map2(lambda x, y: x >= y, val(a[2]), val(a[3]))
            elif n == "FOrdLessThan":
                vals[a[1]] = _map2(lambda x, y: bool(x < y), val(a[2]), val(a[3]))
            elif n == "FOrdGreaterThan":
                vals[a[1]] = _map2(lambda x, y: bool(x > y), val(a[2]), val(a[3]))
            elif n == "FOrdLessThanEqual":
                vals[a[1]] = _map2(lambda x, y: bool(x <= y), val(a[2]), val(a[3]))
            elif n == "FOrdGreaterThanEqual":
                vals[a[1]] = _map2(lambda x, y: bool(x >= y), val(a[2]), val(a[3]))
            elif n == "FOrdEqual":
                vals[a[1]] = _map2(lambda x, y: bool(x == y), val(a[2]), val(a[3]))
            elif n == "LogicalAnd":
                vals[a[1]] = _map2(lambda x, y: bool(x and y), val(a[2]), val(a[3]))
            elif n == "LogicalOr":
                vals[a[1]] = _map2(lambda x, y: bool(x or y), val(a[2]), val(a[3]))
            elif n == "LogicalNot":
                vals[a[1]] = _map1(lambda x: not x, val(a[2]))
            elif n == "Select":
                c, x, y = val(a[2]), val(a[3]), val(a[4])
                vals[a[1]] = [xx if cc else yy for cc, xx, yy in zip(c, x, y)] if isinstance(c, list) else (x if c else y)
            elif n == "ConvertUToF":
                vals[a[1]] = _map1(lambda x: f32(int(x)), val(a[2]))
            elif n == "UConvert":
                vals[a[1]] = _map1(lambda x: int(x) & self._mask(a[0]), val(a[2]))
            elif n == "Bitcast":
                vals[a[1]] = self._bitcast(a[0], val(a[2]))
            elif n == "ExtInst":
                assert a[2] == m.ext_glsl
                vals[a[1]] = self._glsl(a[3], [val(x) for x in a[4:]])
            elif n == "AtomicIAdd":
                p = val(a[2])
                old = self._load(p)
                self._store(p, (old + val(a[5])) & 0xFFFFFFFF)
                vals[a[1]] = old
            elif n == "ImageSampleExplicitLod":
                smp, coord = val(a[2]), val(a[3])
                operands = a[4]
                assert operands & 0x2, "only the Lod image operand is supported"
                lod = val(a[5])
                d = smp.fn(coord[0], coord[1], lod)
                vals[a[1]] = [f32(d), f32(0), f32(0), f32(1)]
            elif n == "Undef":
                vals[a[1]] = self._default(a[0])
            elif n == "Switch":
                sel = int(val(a[0]))
                target = a[1]
                for k in range(2, len(a), 2):
                    if a[k] == sel:
                        target = a[k + 1]
                pc = labels[target]
            elif n == "VectorTimesMatrix":      # r_c = dot(v, column c); same left-to-right accumulation as MatrixTimesVector
                v, mat = val(a[2]), val(a[3])
                out = []
                for c in range(len(mat)):
                    acc = f32(mat[c][0] * v[0])
                    for r in range(1, len(v)):
                        acc = f32(acc + f32(mat[c][r] * v[r]))
                    out.append(acc)
                vals[a[1]] = out
            elif n == "ConvertFToU":
                vals[a[1]] = _map1(_ftou, val(a[2]))
            elif n == "ShiftRightLogical":
                vals[a[1]] = _map2(lambda x, y: (int(x) & self._mask(a[0])) >> (int(y) & 31), val(a[2]), val(a[3]))
            elif n == "ShiftLeftLogical":
                vals[a[1]] = _map2(lambda x, y: (int(x) << (int(y) & 31)) & self._mask(a[0]), val(a[2]), val(a[3]))
            elif n == "BitwiseAnd":
                vals[a[1]] = _map2(lambda x, y: int(x) & int(y), val(a[2]), val(a[3]))
            elif n == "BitwiseOr":
                vals[a[1]] = _map2(lambda x, y: int(x) | int(y), val(a[2]), val(a[3]))
            elif n == "ImageFetch":
                img, coord = val(a[2]), val(a[3])
                lod = 0
                if len(a) > 4:
                    assert a[4] == 0x2, "only the Lod image operand is supported"
                    lod = _s32(val(a[5]))
                d = img.fetch(_s32(coord[0]), _s32(coord[1]), lod)
                vals[a[1]] = [f32(d), f32(0), f32(0), f32(1)]
            elif n == "ImageTexelPointer":
                img = self._load(val(a[2]))
                vals[a[1]] = MemPtr(img.bytes, 4 * int(val(a[3])), T[a[0]].pointee)
            elif n == "ImageWrite" and isinstance(val(a[0]), TexelBufferU32):
                img, coord, texel = val(a[0]), int(val(a[1])), val(a[2])
                if 0 <= 4 * coord < len(img.bytes):
                    np.frombuffer(img.bytes, dtype="<u4", count=1, offset=4 * coord)[0] = int(texel[0] if isinstance(texel, list) else texel) & 0xFFFFFFFF
            elif n == "ImageWrite":
                img, coord, texel = val(a[0]), val(a[1]), val(a[2])
                arr = img.array
                x, y = _s32(coord[0]), _s32(coord[1])
                if 0 <= x < arr.shape[1] and 0 <= y < arr.shape[0]:   # out-of-bounds image stores are discarded
                    arr[y, x] = texel[0]
            else:
                raise NotImplementedError("opcode %s (%d)" % (n, op))

    def _mask(self, tid):
        t = self.m.types[tid]
        if t.kind == "vector":
            t = self.m.types[t.elem]
        return (1 << t.width) - 1

    def _bitcast(self, tid, v):
        t = self.m.types[tid]
        if t.kind == "vector" and isinstance(v, list) and len(v) == t.count:   # e.g. uvec2 -> ivec2: component-wise
            return [self._bitcast(t.elem, x) for x in v]
        if t.kind == "pointer":
            if isinstance(v, MemPtr):
                return MemPtr(v.buf, v.off, t.pointee)
            addr = v if not isinstance(v, list) else (int(v[0]) | (int(v[1]) << 32))
            b, o = self._resolve(int(addr))
            return MemPtr(b, o, t.pointee)
        if t.kind == "float":
            return f32(struct.unpack("<f", struct.pack("<I", int(v) & 0xFFFFFFFF))[0])
        if t.kind == "int":
            if isinstance(v, np.floating):
                return struct.unpack("<I", struct.pack("<f", float(v)))[0]
            return int(v) & self._mask(tid)
        raise NotImplementedError("bitcast to " + t.kind)

    def _access(self, base, idx):
        m = self.m
        if isinstance(base, VarPtr):
            return VarPtr(base.cell, base.path + tuple(int(i) for i in idx))
        buf, off, tid, mstride = base.buf, base.off, base.tid, base.mstride
        for i in idx:
            i = int(i)
            t = m.types[tid]
            if t.kind == "struct":
                d = m.member_decor.get((tid, i), {})
                off += d.get(DEC_OFFSET, 0)
                mstride = _mstride(d)
                tid = t.members[i]
            elif t.kind in ("array", "rtarray"):
                off += m.decor[tid][DEC_ARRAY_STRIDE] * i
                tid = t.elem
            elif t.kind == "vector":
                off += m.type_size(t.elem) * i
                tid = t.elem
            elif t.kind == "matrix":
                if isinstance(mstride, tuple):
                    raise NotImplementedError("access chain into a RowMajor matrix")
                off += (mstride or 16) * i
                tid = t.col
            elif t.kind == "pointer":
                raise NotImplementedError("access chain through a pointer member needs a load")
            else:
                raise NotImplementedError("access chain into " + t.kind)
        return MemPtr(buf, off, tid, None, mstride)

    def _load(self, p):
        if isinstance(p, VarPtr):
            v = p.cell[0]
            for k in p.path:
                v = v[k]
            return _deep(v)
        return self._load_mem(p.buf, p.off, p.tid, p.mstride)

    def _store(self, p, v):
        if isinstance(p, VarPtr):
            if not p.path:
                p.cell[0] = _deep(v)
                return
            c = p.cell[0]
            for k in p.path[:-1]:
                c = c[k]
            c[p.path[-1]] = _deep(v)
            return
        self._store_mem(p.buf, p.off, p.tid, v)

    def _glsl(self, fn, x):
        if fn == GLSL_FABS:
            return _map1(lambda v: f32(abs(v)), x[0])
        if fn == GLSL_FLOOR:
            return _map1(lambda v: f32(np.floor(v)), x[0])
        if fn == GLSL_SQRT:
            return _map1(lambda v: f32(np.sqrt(f32(v))), x[0])
        if fn == GLSL_LOG2:
            return _map1(self._log2, x[0])
        if fn == GLSL_FMAX:   # GLSL max(x, y) = y if x < y else x
            return _map2(lambda p, q: q if p < q else p, x[0], x[1])
        if fn == GLSL_FMIN:
            return _map2(lambda p, q: q if q < p else p, x[0], x[1])
        if fn == GLSL_UMAX:
            return _map2(lambda p, q: max(int(p), int(q)), x[0], x[1])
        if fn == GLSL_UMIN:
            return _map2(lambda p, q: min(int(p), int(q)), x[0], x[1])
        if fn == GLSL_LENGTH:
            v = x[0]
            acc = f32(v[0] * v[0])
            for k in range(1, len(v)):
                acc = f32(acc + f32(v[k] * v[k]))
            return f32(np.sqrt(acc))
        if fn == GLSL_CROSS:
            p, q = x
            return [f32(f32(p[1] * q[2]) - f32(q[1] * p[2])), f32(f32(p[2] * q[0]) - f32(q[2] * p[0])), f32(f32(p[0] * q[1]) - f32(q[0] * p[1]))]
        raise NotImplementedError("GLSL.std.450 instruction %d" % fn)

    def _log2(self, v):
        v = float(v)
        if v != v or v < 0:
            return f32(np.nan)
        if v == 0:
            return f32(-np.inf)
        if math.isinf(v):
            return f32(np.inf)
        if self.m.log2_mode == "exact_floor":   # a value whose floor() equals floor(real log2): mantissa/exponent split
            mant, e = math.frexp(v)              # v = mant * 2^e, mant in [0.5, 1)
            return f32((e - 1) + (0.5 if mant > 0.5 else 0.0))
        return f32(math.log2(v))                 # correctly rounded binary32 of the (double) log2


def _mstride(d):
    ms = d.get(DEC_MATRIX_STRIDE)
    return (ms, True) if (ms is not None and d.get(DEC_ROW_MAJOR)) else ms


def _ftou(v):
    """D3D `ftou`: truncate, NaN and negatives -> 0, saturate at 2^32 - 1."""
    v = float(v)
    if not (v > 0.0):
        return 0
    if v >= 4294967296.0:
        return 0xFFFFFFFF
    return int(v)


def _s32(x):
    x = int(x) & 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def _deep(v):
    return [_deep(x) for x in v] if isinstance(v, list) else v


def _map1(f, x):
    return [f(v) for v in x] if isinstance(x, list) else f(x)


def _map2(f, x, y):
    return [f(p, q) for p, q in zip(x, y)] if isinstance(x, list) else f(x, y)
