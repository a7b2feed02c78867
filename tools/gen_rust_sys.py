"""Generates obs-rvc_b200/rust/rvc-cuda-sys/src/lib.rs - the raw FFI mirror of include/rvc_b200.h - from the header, so the
mirror is complete by construction (tests/test_host_cpu.py checks every declared symbol is mirrored).  No Rust toolchain
exists in the build image: the crate is the binding a maintainer adds to the reference workspace."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "rvc_b200.h")
OUT = os.path.join(ROOT, "obs-rvc_b200", "rust", "rvc-cuda-sys", "src", "lib.rs")

BASE = {"float": "c_float", "double": "f64", "int": "c_int", "size_t": "usize", "uint32_t": "u32", "int32_t": "i32",
        "uint64_t": "u64", "int64_t": "i64", "char": "c_char", "void": "c_void", "rvc_ctx": "rvc_ctx",
        "rvc_config": "rvc_config", "rvc_stream_config": "rvc_stream_config", "long long": "i64"}


def rust_type(c: str) -> str:
    """`const float* const*` -> `*const *const c_float` (pointer levels right to left, constness of the pointee)."""
    toks = c.replace("*", " * ").split()
    base, i, const_base = None, 0, False
    words = []
    while i < len(toks) and toks[i] != "*":
        if toks[i] == "const":
            const_base = True
        else:
            words.append(toks[i])
        i += 1
    base = BASE[" ".join(words)]
    levels = []            # constness of what each '*' points to, innermost first
    pointee_const = const_base
    while i < len(toks):
        assert toks[i] == "*"
        levels.append(pointee_const)
        i += 1
        pointee_const = False
        if i < len(toks) and toks[i] == "const":
            pointee_const = True
            i += 1
    t = base
    for is_const in levels:
        t = ("*const " if is_const else "*mut ") + t
    return t


def prototypes(text: str):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for ret, name, args in re.findall(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(rvc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S):
        params = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
                params.append((m.group(2), m.group(1).strip()))
        yield ret.strip(), name, params


def enum_consts(text: str):
    for body in re.findall(r"enum\s+rvc_[a-z_]+\s*\{(.*?)\}", re.sub(r"/\*.*?\*/", "", text, flags=re.S), flags=re.S):
        for name, val in re.findall(r"(RVC_[A-Z0-9_]+)\s*=\s*(-?\d+)", body):
            yield name, int(val)


def generate() -> str:
    text = open(HDR).read()
    out = ["//! Raw FFI of `librvc_b200.so` - GENERATED from `include/rvc_b200.h` by tools/gen_rust_sys.py; do not edit.",
           "//! NOT compiled in the build image (no Rust toolchain); the crate a maintainer adds to the reference workspace.",
           "#![allow(non_camel_case_types)]",
           "use std::os::raw::{c_char, c_float, c_int, c_void};", "",
           "#[repr(C)]", "pub struct rvc_ctx { _private: [u8; 0] }", ""]
    for name, val in enum_consts(text):
        out.append(f"pub const {name}: c_int = {val};")
    out += ["", "#[repr(C)]", "#[derive(Clone, Copy)]", "pub struct rvc_config {",
            "    pub device: i32,", "    pub noise_mode: i32,", "    pub noise_seed: u64,", "    pub index_k: i32,",
            "    pub upstream_pitch_shift: i32,", "    pub upstream_cents_window: i32,", "    pub use_cuda_graph: i32,",
            "    pub debug_keep: i32,", "    pub reserved: [i32; 7],", "}", "",
            "#[repr(C)]", "#[derive(Clone, Copy)]", "pub struct rvc_stream_config {",
            "    pub sample_rate: u32,", "    pub pitch_shift: i32,", "    pub sample_length: f64,", "    pub crossfade_length: f64,",
            "    pub extra_inference_time: f64,", "    pub rms_mix_rate: f64,", "    pub skip_inference: i32,",
            "    pub reserved: [i32; 7],", "}", "", 'extern "C" {']
    for ret, name, params in prototypes(text):
        ps = ", ".join(f"{('r#in' if p == 'in' else p)}: {rust_type(t)}" for p, t in params)
        r = "" if ret == "void" else f" -> {rust_type(ret)}"
        out.append(f"    pub fn {name}({ps}){r};")
    out += ["}", ""]
    return "\n".join(out)


if __name__ == "__main__":
    src = generate()
    if "--check" in sys.argv:
        sys.exit(0 if open(OUT).read() == src else 1)
    open(OUT, "w").write(src)
    print(f"wrote {OUT}: {src.count('pub fn ')} functions")
