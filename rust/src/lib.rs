//! The public API of ajalab/fm-index 0.3.1 for the count / locate / extract path
//! (src/frontend.rs:26-104, 195-267; src/text.rs; src/error.rs; src/piece.rs), backed by the fmx
//! B200 engine through its C ABI.  Same names, same argument meaning, same results (bit-exact SA
//! ranges, counts, positions in iteration order, piece ids, extracted characters).
//!
//! Source only: the build image has no Rust toolchain, so this crate has not been compiled there.
//! The identical surface is exercised by the Python mirror (fm-index_b200/__init__.py) and the C++
//! mirror (include/fmx.hpp) in tests/test_gpu_parity.py.
pub mod ffi;

use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::c_int;

/// src/error.rs:1-20
#[derive(Debug)]
pub enum Error {
    InvalidText(String),
}
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        match self {
            Error::InvalidText(msg) => write!(f, "invalid text: {}", msg),
        }
    }
}
impl std::error::Error for Error {}

/// src/piece.rs
#[derive(Clone, Copy, Debug, Eq, PartialEq, PartialOrd, Ord, Hash)]
pub struct PieceId(usize);
impl From<usize> for PieceId {
    fn from(v: usize) -> Self { PieceId(v) }
}
impl From<PieceId> for usize {
    fn from(v: PieceId) -> usize { v.0 }
}

/// src/text.rs:10-64 (u8 characters: every BASELINE config)
pub struct Text<T: AsRef<[u8]>> {
    text: T,
    max_character: u8,
}
impl<T: AsRef<[u8]>> Text<T> {
    pub fn new(text: T) -> Self { Text { text, max_character: u8::MAX } }
    pub fn with_max_character(text: T, max_character: u8) -> Self { Text { text, max_character } }
    pub fn text(&self) -> &[u8] { self.text.as_ref() }
    pub fn max_character(&self) -> u8 { self.max_character }
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::fmx_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: c_int) {
    if rc != 0 {
        // a pattern character above max_character panics in the reference too (fm_index.rs:94)
        panic!("fmx: {}", last_error());
    }
}

/// Owner of the device-resident index (replaces SearchIndexWrapper<Backend>, wrapper.rs:10-12).
struct Device(*mut ffi::fmx_index);
unsafe impl Send for Device {}
unsafe impl Sync for Device {}
impl Drop for Device {
    fn drop(&mut self) { unsafe { ffi::fmx_index_free(self.0) } }
}
impl Device {
    fn build<T: AsRef<[u8]>>(text: &Text<T>, kind: c_int, level: c_int) -> Result<Self, Error> {
        let t = text.text();
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            ffi::fmx_index_build(t.as_ptr().cast(), t.len() as u64, 1, text.max_character() as u64, kind, level, 0, &mut h)
        };
        match rc {
            0 => Ok(Device(h)),
            ffi::FMX_ERR_INVALID_TEXT => Err(Error::InvalidText(last_error())),
            _ => panic!("fmx: {}", last_error()),
        }
    }
}

/// Search result (frontend.rs:65-83 over wrapper.rs:14-23).
pub struct Search<'a> {
    dev: &'a Device,
    mode: c_int,
    s: usize,
    e: usize,
    fresh: bool,
    locate: bool,
    multi: bool,
}
impl<'a> Search<'a> {
    /// Refine: prepends `pattern` (wrapper.rs:99-124).
    pub fn search<K: AsRef<[u8]>>(&self, pattern: K) -> Self {
        let p = pattern.as_ref();
        let off = [0u64, p.len() as u64];
        let (is, ie) = (self.s as u64, self.e as u64);
        let (mut s, mut e) = (0u64, 0u64);
        let (pis, pie) = if self.fresh { (std::ptr::null(), std::ptr::null()) } else { (&is as *const u64, &ie as *const u64) };
        check(unsafe { ffi::fmx_search_batch(self.dev.0, self.mode, p.as_ptr(), off.as_ptr(), 0, 1, pis, pie, &mut s, &mut e) });
        Search { dev: self.dev, mode: self.mode, s: s as usize, e: e as usize, fresh: false, locate: self.locate, multi: self.multi }
    }
    /// wrapper.rs:132-134
    pub fn count(&self) -> usize { self.e - self.s }
    /// wrapper.rs:137-139, 203-217: ascending SA-row order, `L == 0` filter for prefix/exact searches.
    pub fn iter_matches(&'a self) -> impl Iterator<Item = Match<'a>> + 'a {
        let prefix_only = self.mode == ffi::FMX_SEARCH_PREFIX || self.mode == ffi::FMX_SEARCH_EXACT;
        let mut rows: Vec<u64> = (self.s as u64..self.e as u64).collect();
        if prefix_only && !rows.is_empty() {
            let mut l = vec![0u64; rows.len()];
            check(unsafe { ffi::fmx_rows_op(self.dev.0, 0, rows.as_ptr(), rows.len() as u64, l.as_mut_ptr()) });
            rows = rows.into_iter().zip(l).filter(|(_, c)| *c == 0).map(|(r, _)| r).collect();
        }
        let (mut pos, mut pid): (Vec<u64>, Vec<u64>) = (vec![], vec![]);
        if self.locate && !rows.is_empty() {
            let (s, e) = (self.s as u64, self.e as u64);
            let mut off = [0u64; 2];
            let (mut p, mut d) = (std::ptr::null_mut(), std::ptr::null_mut());
            let pd = if self.multi { &mut d as *mut *mut u64 } else { std::ptr::null_mut() };
            check(unsafe { ffi::fmx_locate_batch(self.dev.0, prefix_only as c_int, &s, &e, 1, off.as_mut_ptr(), &mut p, pd) });
            unsafe {
                if !p.is_null() { pos = std::slice::from_raw_parts(p, off[1] as usize).to_vec(); ffi::fmx_free(p.cast()); }
                if !d.is_null() { pid = std::slice::from_raw_parts(d, off[1] as usize).to_vec(); ffi::fmx_free(d.cast()); }
            }
        }
        let dev = self.dev;
        rows.into_iter().enumerate().map(move |(k, i)| Match {
            dev,
            i: i as usize,
            position: pos.get(k).map(|v| *v as usize),
            piece: pid.get(k).map(|v| PieceId(*v as usize)),
        })
    }
    /// Text positions of the matches in pages of at most `page` entries, in `iter_matches()` order, without
    /// ever holding the whole list (fmx_locate_page): the bounded-memory form of the lazy iterator for
    /// searches with very many matches.  Unfiltered searches only (search / search_suffix).
    pub fn locate_pages(&'a self, page: usize) -> impl Iterator<Item = Vec<u64>> + 'a {
        assert!(self.locate, "locate_pages() needs a ...WithLocate index");
        assert!(self.mode == ffi::FMX_SEARCH || self.mode == ffi::FMX_SEARCH_SUFFIX, "paged locate is unfiltered");
        let (s, e, dev) = (self.s as u64, self.e as u64, self.dev);
        let mut first = 0u64;
        let mut done = page == 0;
        std::iter::from_fn(move || {
            if done { return None; }
            let mut buf = vec![0u64; page];
            let mut total = 0u64;
            check(unsafe { ffi::fmx_locate_page(dev.0, &s, &e, 1, first, page as u64, buf.as_mut_ptr(), std::ptr::null_mut(), &mut total) });
            let got = total.saturating_sub(first).min(page as u64) as usize;
            first += page as u64;
            done = first >= total;
            if got == 0 { return None; }
            buf.truncate(got);
            Some(buf)
        })
    }
}

/// Match / MatchWithLocate / MatchWithPieceId (frontend.rs:85-104 over wrapper.rs:219-248).
pub struct Match<'a> {
    dev: &'a Device,
    i: usize,
    position: Option<usize>,
    piece: Option<PieceId>,
}
impl<'a> Match<'a> {
    /// MatchWithLocate::locate
    pub fn locate(&self) -> usize { self.position.expect("locate() needs a ...WithLocate index") }
    /// MatchWithPieceId::piece_id
    pub fn piece_id(&self) -> PieceId { self.piece.expect("piece_id() needs FMIndexMultiPiecesWithLocate") }
    fn chars(&self, forward: bool) -> impl Iterator<Item = u8> + 'a {
        // lazily pulls blocks of characters; backward never ends (the walk is cyclic, wrapper.rs:154-161),
        // forward ends where fl_map is None (multi_pieces.rs:171-181)
        let (dev, row) = (self.dev, self.i as u64);
        let mut buf: Vec<u8> = vec![];
        let (mut served, mut want, mut ended) = (0usize, 32u32, false);
        std::iter::from_fn(move || {
            if served == buf.len() && !ended {
                want *= 2;
                buf = vec![0u8; want as usize];
                let mut got = 0u32;
                check(unsafe { ffi::fmx_extract_batch(dev.0, &row, 1, want, forward as c_int, buf.as_mut_ptr(), &mut got) });
                buf.truncate(got as usize);
                ended = got < want;
            }
            if served < buf.len() { served += 1; Some(buf[served - 1]) } else { None }
        })
    }
    pub fn iter_chars_forward(&self) -> impl Iterator<Item = u8> + 'a { self.chars(true) }
    pub fn iter_chars_backward(&self) -> impl Iterator<Item = u8> + 'a { self.chars(false) }
}

macro_rules! index_type {
    ($name:ident, $kind:expr, count_only) => {
        pub struct $name(Device, PhantomData<u8>);
        impl $name {
            /// frontend.rs:195-203 / 220-228 / 245-252
            pub fn new<T: AsRef<[u8]>>(text: &Text<T>) -> Result<Self, Error> {
                Ok($name(Device::build(text, $kind, ffi::FMX_LEVEL_COUNT_ONLY)?, PhantomData))
            }
            index_type!(@common $kind, false);
        }
    };
    ($name:ident, $kind:expr, with_locate) => {
        pub struct $name(Device, PhantomData<u8>);
        impl $name {
            /// frontend.rs:205-218 / 230-243 / 254-267
            pub fn new<T: AsRef<[u8]>>(text: &Text<T>, level: usize) -> Result<Self, Error> {
                Ok($name(Device::build(text, $kind, level as c_int)?, PhantomData))
            }
            index_type!(@common $kind, true);
        }
    };
    (@common $kind:expr, $locate:expr) => {
        fn start(&self, mode: c_int) -> Search<'_> {
            Search { dev: &self.0, mode, s: 0, e: 0, fresh: true, locate: $locate, multi: $kind == ffi::FMX_KIND_MULTI }
        }
        /// SearchIndex::search (frontend.rs:291-296)
        pub fn search<K: AsRef<[u8]>>(&self, pattern: K) -> Search<'_> { self.start(ffi::FMX_SEARCH).search(pattern) }
        /// SearchIndex::len: includes the trailing \0 (frontend.rs:35-39)
        pub fn len(&self) -> usize { unsafe { ffi::fmx_index_len(self.0 .0) as usize } }
        /// SearchIndex::heap_size: bytes of the device-resident index (frontend.rs:41-44)
        pub fn heap_size(&self) -> usize { unsafe { ffi::fmx_index_device_bytes(self.0 .0) as usize } }
        /// Batched search + locate over many patterns: (SA ranges, CSR hit offsets, positions).
        pub fn search_locate_batch<K: AsRef<[u8]>>(&self, patterns: &[K]) -> (Vec<(usize, usize)>, Vec<u64>, Vec<u64>) {
            let mut flat: Vec<u8> = vec![];
            let mut off: Vec<u64> = vec![0];
            for p in patterns { flat.extend_from_slice(p.as_ref()); off.push(flat.len() as u64); }
            let n = patterns.len();
            let (mut s, mut e, mut hoff) = (vec![0u64; n], vec![0u64; n], vec![0u64; n + 1]);
            check(unsafe { ffi::fmx_search_batch(self.0 .0, ffi::FMX_SEARCH, flat.as_ptr(), off.as_ptr(), 0, n as u64,
                                                 std::ptr::null(), std::ptr::null(), s.as_mut_ptr(), e.as_mut_ptr()) });
            let mut pos: Vec<u64> = vec![];
            if $locate {
                let mut p = std::ptr::null_mut();
                check(unsafe { ffi::fmx_locate_batch(self.0 .0, 0, s.as_ptr(), e.as_ptr(), n as u64, hoff.as_mut_ptr(), &mut p,
                                                     std::ptr::null_mut()) });
                unsafe { if !p.is_null() { pos = std::slice::from_raw_parts(p, hoff[n] as usize).to_vec(); ffi::fmx_free(p.cast()); } }
            }
            (s.iter().zip(&e).map(|(a, b)| (*a as usize, *b as usize)).collect(), hoff, pos)
        }
    };
}

macro_rules! multi_pieces_searches {
    ($name:ident) => {
        /// SearchIndexWithMultiPieces (frontend.rs:47-63, 369-390)
        impl $name {
            pub fn search_prefix<K: AsRef<[u8]>>(&self, p: K) -> Search<'_> { self.start(ffi::FMX_SEARCH_PREFIX).search(p) }
            pub fn search_suffix<K: AsRef<[u8]>>(&self, p: K) -> Search<'_> { self.start(ffi::FMX_SEARCH_SUFFIX).search(p) }
            pub fn search_exact<K: AsRef<[u8]>>(&self, p: K) -> Search<'_> { self.start(ffi::FMX_SEARCH_EXACT).search(p) }
            pub fn pieces_count(&self) -> usize { unsafe { ffi::fmx_index_pieces_count(self.0 .0) as usize } }
        }
    };
}

index_type!(FMIndex, ffi::FMX_KIND_FM, count_only);
index_type!(FMIndexWithLocate, ffi::FMX_KIND_FM, with_locate);
index_type!(RLFMIndex, ffi::FMX_KIND_RLFM, count_only);
index_type!(RLFMIndexWithLocate, ffi::FMX_KIND_RLFM, with_locate);
index_type!(FMIndexMultiPieces, ffi::FMX_KIND_MULTI, count_only);
index_type!(FMIndexMultiPiecesWithLocate, ffi::FMX_KIND_MULTI, with_locate);
multi_pieces_searches!(FMIndexMultiPieces);
multi_pieces_searches!(FMIndexMultiPiecesWithLocate);
