//! The public API of ajalab/fm-index 0.3.1 for the count / locate / extract path, backed by the fmx B200 engine
//! through its C ABI (include/fmx.h).
//!
//! Trait for trait and type for type what `src/lib.rs:134-146` of the reference exports:
//! `Character`, `Error`, `PieceId`, `Text`, the traits `SearchIndex`, `SearchIndexWithMultiPieces`, `Search`, `Match`,
//! `MatchWithLocate`, `MatchWithPieceId` (`src/frontend.rs:26-104`) and the eighteen concrete types
//! (`FMIndex` .. `FMIndexMultiPiecesMatchWithLocate`), all generic over `C: Character`, so code written against the
//! traits -- e.g. the reference's own `tests/test_api.rs:5-11`, `fn len<T: SearchIndex<u8>>(index: &T)` -- compiles
//! unchanged.  Results are bit-exact: SA ranges, counts, positions in iteration order, piece ids, extracted characters.
//!
//! The structure is NOT the reference's (one wrapper + one backend type per index, stitched by macros): here one
//! `Core` owns the device-resident index, and the search / match types are two generic structs tagged by a
//! zero-sized capability marker.  Beyond the reference: `search_batch` (many patterns, one fused GPU call).
//!
//! Source only: the build image has no Rust toolchain, so this crate has not been compiled there.  The identical
//! surface is exercised by the Python mirror (fm-index_b200/__init__.py) and the C++ mirror (include/fmx.hpp) in
//! tests/test_gpu_parity.py.
pub mod ffi;

use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::c_int;

use num_traits::Bounded;

// ------------------------------------------------------------------ Character, Error, PieceId, Text

/// `src/character.rs:6-42`
pub trait Character: Copy + Clone {
    fn into_u64(self) -> u64;
    fn from_u64(x: u64) -> Self;
    fn into_usize(self) -> usize {
        self.into_u64() as usize
    }
    fn from_usize(x: usize) -> Self {
        Self::from_u64(x as u64)
    }
}
macro_rules! character {
    ($($t:ty),*) => {$(
        impl Character for $t {
            fn into_u64(self) -> u64 { self as u64 }
            fn from_u64(x: u64) -> Self { x as $t }
        }
    )*};
}
character!(u8, u16, u32, u64, usize);

/// `src/error.rs:1-20`
#[derive(Debug)]
pub enum Error {
    InvalidText(&'static str),
}
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        match self {
            Error::InvalidText(msg) => write!(f, "invalid text: {}", msg),
        }
    }
}
impl std::error::Error for Error {}

/// `src/piece.rs`
#[derive(Clone, Copy, Debug, Eq, PartialEq, PartialOrd, Ord, Hash)]
pub struct PieceId(usize);
impl From<usize> for PieceId {
    fn from(v: usize) -> Self {
        PieceId(v)
    }
}
impl From<PieceId> for usize {
    fn from(v: PieceId) -> usize {
        v.0
    }
}

/// `src/text.rs:10-64`
pub struct Text<C, T>
where
    C: Character,
    T: AsRef<[C]>,
{
    text: T,
    max_character: C,
}
impl<C, T> Text<C, T>
where
    C: Character + Bounded,
    T: AsRef<[C]>,
{
    pub fn new(text: T) -> Self {
        Text { text, max_character: C::max_value() }
    }
}
impl<C, T> Text<C, T>
where
    C: Character,
    T: AsRef<[C]>,
{
    pub fn with_max_character(text: T, max_character: C) -> Self {
        Text { text, max_character }
    }
    pub fn text(&self) -> &[C] {
        self.text.as_ref()
    }
    pub fn max_character(&self) -> C {
        self.max_character
    }
}

// ------------------------------------------------------------------ the traits (src/frontend.rs:26-104)

pub trait SearchIndex<C> {
    fn search<K>(&self, pattern: K) -> impl Search<'_, C>
    where
        K: AsRef<[C]>;
    /// includes the trailing `\0`
    fn len(&self) -> usize;
    /// here: bytes of the device-resident index
    fn heap_size(&self) -> usize;
}

pub trait SearchIndexWithMultiPieces<C>: SearchIndex<C> {
    fn search_prefix<K>(&self, pattern: K) -> impl Search<'_, C>
    where
        K: AsRef<[C]>;
    fn search_suffix<K>(&self, pattern: K) -> impl Search<'_, C>
    where
        K: AsRef<[C]>;
    fn search_exact<K>(&self, pattern: K) -> impl Search<'_, C>
    where
        K: AsRef<[C]>;
}

pub trait Search<'a, C> {
    type Match: Match<'a, C>;
    fn search<K: AsRef<[C]>>(&self, pattern: K) -> Self;
    fn count(&self) -> usize;
    fn iter_matches(&'a self) -> impl Iterator<Item = Self::Match> + 'a;
}

pub trait Match<'a, C> {
    fn iter_chars_forward(&self) -> impl Iterator<Item = C> + 'a;
    fn iter_chars_backward(&self) -> impl Iterator<Item = C> + 'a;
}

pub trait MatchWithLocate<'a, C>: Match<'a, C> {
    fn locate(&self) -> usize;
}

pub trait MatchWithPieceId<'a, C>: Match<'a, C> {
    fn piece_id(&self) -> PieceId;
}

// ------------------------------------------------------------------ the engine handle

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::fmx_last_error()).to_string_lossy().into_owned() }
}

/// Queries are infallible in the reference; the one way they fail there is a pattern character above
/// `max_character`, which panics (index out of bounds on `cs`, `src/fm_index.rs:94`).  Same here.
fn check(rc: c_int) {
    if rc != 0 {
        panic!("fmx: {}", last_error());
    }
}

/// the three messages of `src/suffix_array/sais.rs:128-139`, as the `&'static str` the reference's `Error` carries
fn invalid_text(msg: &str) -> Error {
    if msg.contains("must not start") {
        Error::InvalidText("the given text must not start with zero character")
    } else if msg.contains("exactly one zero") {
        Error::InvalidText("the given text must end with exactly one zero character")
    } else {
        Error::InvalidText("the given text cannot be indexed")
    }
}

/// Owner of one device-resident index.  Immutable after construction: `Send + Sync` like the reference's indexes.
struct Core {
    h: *mut ffi::fmx_index,
}
unsafe impl Send for Core {}
unsafe impl Sync for Core {}
impl Drop for Core {
    fn drop(&mut self) {
        unsafe { ffi::fmx_index_free(self.h) }
    }
}

impl Core {
    fn build<C: Character, T: AsRef<[C]>>(text: &Text<C, T>, kind: c_int, level: c_int) -> Result<Core, Error> {
        let t = text.text();
        let mut h = std::ptr::null_mut();
        // characters travel in their own width; the library says FMX_ERR_UNSUPPORTED for widths it has no layout for
        let rc = unsafe {
            ffi::fmx_index_build(t.as_ptr().cast(), t.len() as u64, std::mem::size_of::<C>() as u32, text.max_character().into_u64(),
                                 kind, level, 0, &mut h)
        };
        match rc {
            0 => Ok(Core { h }),
            ffi::FMX_ERR_INVALID_TEXT => Err(invalid_text(&last_error())),
            _ => panic!("fmx: {}", last_error()),
        }
    }

    /// one pattern, fresh or refining `(s, e)` (wrapper.rs:99-124)
    fn search<C: Character>(&self, mode: c_int, pattern: &[C], from: Option<(u64, u64)>) -> (u64, u64) {
        let off = [0u64, pattern.len() as u64];
        let (mut s, mut e) = (0u64, 0u64);
        let (is, ie) = from.unwrap_or((0, 0));
        let (pis, pie) = if from.is_some() { (&is as *const u64, &ie as *const u64) } else { (std::ptr::null(), std::ptr::null()) };
        check(unsafe { ffi::fmx_search_batch(self.h, mode, pattern.as_ptr().cast(), off.as_ptr(), 0, 1, pis, pie, &mut s, &mut e) });
        (s, e)
    }
}

// ------------------------------------------------------------------ capability markers

/// What an index type can do, as zero-sized markers: the reference spells this with six backend instantiations.
pub trait Flavor {
    const KIND: c_int;
    const LOCATE: bool;
}
pub trait CanLocate: Flavor {}
macro_rules! flavor {
    ($name:ident, $kind:expr, $locate:expr) => {
        #[doc(hidden)]
        pub struct $name;
        impl Flavor for $name {
            const KIND: c_int = $kind;
            const LOCATE: bool = $locate;
        }
    };
}
flavor!(FmCount, ffi::FMX_KIND_FM, false);
flavor!(FmLocate, ffi::FMX_KIND_FM, true);
flavor!(RlCount, ffi::FMX_KIND_RLFM, false);
flavor!(RlLocate, ffi::FMX_KIND_RLFM, true);
flavor!(MultiCount, ffi::FMX_KIND_MULTI, false);
flavor!(MultiLocate, ffi::FMX_KIND_MULTI, true);
impl CanLocate for FmLocate {}
impl CanLocate for RlLocate {}
impl CanLocate for MultiLocate {}

// ------------------------------------------------------------------ Search and Match, once, for every flavor

/// The result of a search (`Search` trait); borrows the index.
pub struct SearchOf<'a, C: Character, F: Flavor> {
    core: &'a Core,
    mode: c_int,
    range: (u64, u64),
    _m: PhantomData<(C, F)>,
}

/// One match: an SA row, with its text position / piece id fetched by the batch call of `iter_matches`.
pub struct MatchOf<'a, C: Character, F: Flavor> {
    core: &'a Core,
    row: u64,
    position: Option<u64>,
    piece: Option<u64>,
    _m: PhantomData<(C, F)>,
}

impl<'a, C: Character, F: Flavor> SearchOf<'a, C, F> {
    fn fresh<K: AsRef<[C]>>(core: &'a Core, mode: c_int, pattern: K) -> Self {
        SearchOf { core, mode, range: core.search(mode, pattern.as_ref(), None), _m: PhantomData }
    }
    fn prefix_only(&self) -> bool {
        self.mode == ffi::FMX_SEARCH_PREFIX || self.mode == ffi::FMX_SEARCH_EXACT
    }
    /// Refine: prepends `pattern` (inherent twin of `Search::search`, as in the reference).
    pub fn search<K: AsRef<[C]>>(&self, pattern: K) -> Self {
        SearchOf { core: self.core, mode: self.mode, range: self.core.search(self.mode, pattern.as_ref(), Some(self.range)), _m: PhantomData }
    }
    /// `e - s`, ignoring the prefix filter (wrapper.rs:132-134)
    pub fn count(&self) -> usize {
        (self.range.1 - self.range.0) as usize
    }
}

impl<'a, C: Character, F: Flavor> Search<'a, C> for SearchOf<'a, C, F> {
    type Match = MatchOf<'a, C, F>;

    fn search<K: AsRef<[C]>>(&self, pattern: K) -> Self {
        SearchOf::search(self, pattern)
    }

    fn count(&self) -> usize {
        SearchOf::count(self)
    }

    /// Rows `s..e` ascending, rows with `L != 0` dropped for prefix / exact searches (wrapper.rs:137-139, 203-217).
    /// Lazy: rows, positions and piece ids arrive in pages of 4096 matches (fmx_locate_page), so a search with 10^8
    /// matches never holds more than a page; the filtered modes use one fmx_locate_batch call.
    fn iter_matches(&'a self) -> impl Iterator<Item = Self::Match> + 'a {
        const PAGE: u64 = 4096;
        let (s, e) = self.range;
        let core = self.core;
        let filtered = self.prefix_only();
        let multi = F::KIND == ffi::FMX_KIND_MULTI;
        let mut page: Vec<(u64, Option<u64>, Option<u64>)> = Vec::new();
        let (mut served, mut next_row, mut done) = (0usize, s, e <= s);
        std::iter::from_fn(move || {
            if served == page.len() && !done {
                page.clear();
                served = 0;
                if filtered {
                    // one shot: the kept rows are not a contiguous range
                    let n = (e - s) as usize;
                    let rows: Vec<u64> = (s..e).collect();
                    let mut l = vec![0u64; n];
                    check(unsafe { ffi::fmx_rows_op(core.h, 0, rows.as_ptr(), n as u64, l.as_mut_ptr()) });
                    let kept: Vec<u64> = rows.into_iter().zip(l).filter(|(_, c)| *c == 0).map(|(r, _)| r).collect();
                    let (mut pos, mut pid): (Vec<u64>, Vec<u64>) = (vec![], vec![]);
                    if F::LOCATE && !kept.is_empty() {
                        let mut off = [0u64; 2];
                        let (mut p, mut d) = (std::ptr::null_mut(), std::ptr::null_mut());
                        let pd = if multi { &mut d as *mut *mut u64 } else { std::ptr::null_mut() };
                        check(unsafe { ffi::fmx_locate_batch(core.h, 1, &s, &e, 1, off.as_mut_ptr(), &mut p, pd) });
                        unsafe {
                            if !p.is_null() {
                                pos = std::slice::from_raw_parts(p, off[1] as usize).to_vec();
                                ffi::fmx_free(p.cast());
                            }
                            if !d.is_null() {
                                pid = std::slice::from_raw_parts(d, off[1] as usize).to_vec();
                                ffi::fmx_free(d.cast());
                            }
                        }
                    }
                    for (k, r) in kept.into_iter().enumerate() {
                        page.push((r, pos.get(k).copied(), pid.get(k).copied()));
                    }
                    done = true;
                } else {
                    let n = PAGE.min(e - next_row);
                    let (mut pos, mut pid) = (vec![0u64; n as usize], vec![0u64; n as usize]);
                    if F::LOCATE {
                        let mut total = 0u64;
                        let pp = if multi { pid.as_mut_ptr() } else { std::ptr::null_mut() };
                        check(unsafe { ffi::fmx_locate_page(core.h, &s, &e, 1, next_row - s, n, pos.as_mut_ptr(), pp, &mut total) });
                    }
                    for k in 0..n as usize {
                        page.push((next_row + k as u64, if F::LOCATE { Some(pos[k]) } else { None },
                                   if F::LOCATE && multi { Some(pid[k]) } else { None }));
                    }
                    next_row += n;
                    done = next_row >= e;
                }
            }
            if served < page.len() {
                served += 1;
                let (row, position, piece) = page[served - 1];
                Some(MatchOf { core, row, position, piece, _m: PhantomData })
            } else {
                None
            }
        })
    }
}

impl<'a, C: Character, F: Flavor> MatchOf<'a, C, F> {
    /// Lazily pulls blocks of characters (32, 64, 128 ..): backward never ends (the walk is cyclic, wrapper.rs:154-161),
    /// forward ends where `fl_map` is `None` (multi_pieces.rs:171-181).
    fn chars(&self, forward: bool) -> impl Iterator<Item = C> + 'a {
        let (core, row) = (self.core, self.row);
        let width = std::mem::size_of::<C>();
        let mut buf: Vec<u8> = vec![];
        let (mut have, mut served, mut want, mut ended) = (0usize, 0usize, 16u32, false);
        std::iter::from_fn(move || {
            if served == have && !ended {
                want *= 2;
                buf = vec![0u8; want as usize * width];
                let mut got = 0u32;
                check(unsafe { ffi::fmx_extract_batch(core.h, &row, 1, want, forward as c_int, buf.as_mut_ptr(), &mut got) });
                have = got as usize;
                ended = got < want;
            }
            if served < have {
                let mut v = 0u64;
                for b in 0..width {
                    v |= (buf[served * width + b] as u64) << (8 * b);
                }
                served += 1;
                Some(C::from_u64(v))
            } else {
                None
            }
        })
    }
}

impl<'a, C: Character, F: Flavor> Match<'a, C> for MatchOf<'a, C, F> {
    fn iter_chars_forward(&self) -> impl Iterator<Item = C> + 'a {
        self.chars(true)
    }
    fn iter_chars_backward(&self) -> impl Iterator<Item = C> + 'a {
        self.chars(false)
    }
}

impl<'a, C: Character, F: CanLocate> MatchWithLocate<'a, C> for MatchOf<'a, C, F> {
    fn locate(&self) -> usize {
        match self.position {
            Some(p) => p as usize,
            None => {
                let mut out = 0u64;
                check(unsafe { ffi::fmx_rows_op(self.core.h, 4, &self.row, 1, &mut out) });
                out as usize
            }
        }
    }
}

/// only `FMIndexMultiPiecesMatchWithLocate` has it, as in the reference (frontend.rs:542)
impl<'a, C: Character> MatchWithPieceId<'a, C> for MatchOf<'a, C, MultiLocate> {
    fn piece_id(&self) -> PieceId {
        match self.piece {
            Some(d) => PieceId(d as usize),
            None => {
                let mut out = 0u64;
                check(unsafe { ffi::fmx_rows_op(self.core.h, 5, &self.row, 1, &mut out) });
                PieceId(out as usize)
            }
        }
    }
}

// ------------------------------------------------------------------ batched entry (beyond the reference)

/// Counts and matches of many patterns from ONE fused GPU call (fmx_query_batch): `counts[p]`, and the matches of
/// pattern `p` in `positions[hit_off[p] .. hit_off[p + 1]]`, in the reference's iteration order.
pub struct BatchResult {
    pub counts: Vec<u64>,
    pub hit_off: Vec<u64>,
    pub positions: Vec<u64>,
    pub piece_ids: Vec<u64>,
}

fn query_batch<C: Character, K: AsRef<[C]>>(core: &Core, mode: c_int, patterns: &[K], locate: bool, pieces: bool) -> BatchResult {
    let h = core.h;
    query_batch_with::<C, K, _>(|q, total| unsafe { ffi::fmx_query_batch(h, q, total) }, mode, patterns, locate, pieces, locate)
}

/// `offsets`: `hit_off` is wanted even without positions (a piece-partitioned group derives its counts from it).
fn query_batch_with<C: Character, K: AsRef<[C]>, F: Fn(*const ffi::fmx_query, *mut u64) -> c_int>(
    call: F, mode: c_int, patterns: &[K], locate: bool, pieces: bool, offsets: bool) -> BatchResult {
    let mut flat: Vec<C> = vec![];
    let mut off: Vec<u64> = vec![0];
    for p in patterns {
        flat.extend_from_slice(p.as_ref());
        off.push(flat.len() as u64);
    }
    let n = patterns.len();
    let mut res = BatchResult { counts: vec![0; n], hit_off: vec![0; n + 1], positions: vec![], piece_ids: vec![] };
    let mut cap = if locate { 2 * n + 1024 } else { 0 };
    loop {
        if locate {
            res.positions = vec![0; cap];
            if pieces {
                res.piece_ids = vec![0; cap];
            }
        }
        let q = ffi::fmx_query {
            mode,
            packed_bits: 0,
            patterns: flat.as_ptr().cast(),
            pat_off: off.as_ptr(),
            fixed_len: 0,
            npat: n as u64,
            out_width: 8,
            reserved: 0,
            out_s: std::ptr::null_mut(),
            out_e: std::ptr::null_mut(),
            counts: res.counts.as_mut_ptr().cast(),
            hit_off: if locate || offsets { res.hit_off.as_mut_ptr().cast() } else { std::ptr::null_mut() },
            positions: if locate { res.positions.as_mut_ptr().cast() } else { std::ptr::null_mut() },
            piece_ids: if locate && pieces { res.piece_ids.as_mut_ptr().cast() } else { std::ptr::null_mut() },
            capacity: cap as u64,
        };
        let mut total = 0u64;
        let rc = call(&q, &mut total);
        if rc == ffi::FMX_ERR_CAPACITY {
            cap = total as usize; // hit_off and the total are valid: once more with room for every match
            continue;
        }
        check(rc);
        res.positions.truncate(total as usize);
        res.piece_ids.truncate(total as usize);
        return res;
    }
}

// ------------------------------------------------------------------ several GPUs of this process (beyond the reference)

/// How an [`IndexGroup`] uses its devices.
pub enum GroupMode {
    /// The index copied to every device, batches cut into shards: input order and the reference's iteration order
    /// are preserved.
    Replicate,
    /// `FMIndexMultiPieces` only: the pieces partitioned over the devices (device ids may repeat), every member answers
    /// the whole batch, the parts are merged on the first device.  Counts, match sets, positions and piece ids are
    /// those of the whole text (`multi_pieces.rs:188-223`); the matches of a pattern come partition-major.  This is
    /// also how a text of 2^32 symbols or more is served: every partition stays below 2^32.
    ByPiece,
}

/// One handle over several GPUs (`fmx_group_*`); u8 texts.
pub struct IndexGroup {
    g: *mut ffi::fmx_group,
    locate: bool,
    multi: bool,
}

impl Drop for IndexGroup {
    fn drop(&mut self) {
        unsafe { ffi::fmx_group_free(self.g) }
    }
}

impl IndexGroup {
    /// `kind`: `ffi::FMX_KIND_*`; `level`: the sampling level of the `WithLocate` types, `None` for count-only ones.
    pub fn new<T: AsRef<[u8]>>(text: &Text<u8, T>, kind: c_int, level: Option<usize>, devices: &[c_int], mode: GroupMode)
                               -> Result<IndexGroup, Error> {
        let t = text.text();
        let mut g: *mut ffi::fmx_group = std::ptr::null_mut();
        let gm = match mode { GroupMode::Replicate => ffi::FMX_GROUP_REPLICATE, GroupMode::ByPiece => ffi::FMX_GROUP_BY_PIECE };
        let lvl = level.map(|l| l as c_int).unwrap_or(ffi::FMX_LEVEL_COUNT_ONLY);
        let rc = unsafe {
            ffi::fmx_group_create(devices.as_ptr(), devices.len() as c_int, gm, t.as_ptr().cast(), t.len() as u64, 1,
                                  text.max_character().into_u64(), kind, lvl, ffi::FMX_MODE_AUTO, &mut g)
        };
        match rc {
            0 => Ok(IndexGroup { g, locate: level.is_some(), multi: kind == ffi::FMX_KIND_MULTI }),
            ffi::FMX_ERR_INVALID_TEXT => Err(invalid_text(&last_error())),
            _ => panic!("fmx: {}", last_error()),
        }
    }
    /// The size of the whole text, including the trailing `\0`.
    pub fn len(&self) -> usize {
        unsafe { ffi::fmx_group_len(self.g) as usize }
    }
    pub fn pieces_count(&self) -> usize {
        unsafe { ffi::fmx_group_pieces_count(self.g) as usize }
    }
    /// Count (and locate) many patterns on all devices of the group.
    pub fn search_batch<K: AsRef<[u8]>>(&self, patterns: &[K]) -> BatchResult {
        let g = self.g;
        query_batch_with::<u8, K, _>(|q, total| unsafe { ffi::fmx_group_query_batch(g, q, total) }, ffi::FMX_SEARCH, patterns,
                                     self.locate, self.locate && self.multi, true)
    }
}

// ------------------------------------------------------------------ the six index types and their aliases

macro_rules! index_type {
    ($name:ident, $flavor:ident, $search:ident, $matchty:ident, $doc:expr) => {
        #[doc = $doc]
        pub struct $name<C: Character>(Core, PhantomData<C>);
        pub type $search<'a, C> = SearchOf<'a, C, $flavor>;
        pub type $matchty<'a, C> = MatchOf<'a, C, $flavor>;

        impl<C: Character> SearchIndex<C> for $name<C> {
            fn search<K>(&self, pattern: K) -> impl Search<'_, C>
            where
                K: AsRef<[C]>,
            {
                $name::search(self, pattern)
            }
            fn len(&self) -> usize {
                unsafe { ffi::fmx_index_len(self.0.h) as usize }
            }
            fn heap_size(&self) -> usize {
                unsafe { ffi::fmx_index_device_bytes(self.0.h) as usize }
            }
        }

        impl<C: Character> $name<C> {
            /// Search for a pattern in the text.
            pub fn search<K: AsRef<[C]>>(&self, pattern: K) -> $search<'_, C> {
                SearchOf::fresh(&self.0, ffi::FMX_SEARCH, pattern)
            }
            /// The size of the text in the index, including the trailing `\0`.
            pub fn len(&self) -> usize {
                SearchIndex::len(self)
            }
            /// Count (and, for the `WithLocate` types, locate) many patterns in one fused GPU call.
            pub fn search_batch<K: AsRef<[C]>>(&self, patterns: &[K]) -> BatchResult {
                query_batch::<C, K>(&self.0, ffi::FMX_SEARCH, patterns, <$flavor as Flavor>::LOCATE,
                                    <$flavor as Flavor>::LOCATE && <$flavor as Flavor>::KIND == ffi::FMX_KIND_MULTI)
            }
        }
    };
}

macro_rules! count_only_new {
    ($name:ident, $flavor:ident) => {
        impl<C: Character> $name<C> {
            /// `src/frontend.rs:195-203, 220-228, 245-252`
            pub fn new<T: AsRef<[C]>>(text: &Text<C, T>) -> Result<Self, Error> {
                Ok($name(Core::build(text, <$flavor as Flavor>::KIND, ffi::FMX_LEVEL_COUNT_ONLY)?, PhantomData))
            }
        }
    };
}
macro_rules! with_locate_new {
    ($name:ident, $flavor:ident) => {
        impl<C: Character> $name<C> {
            /// `src/frontend.rs:205-218, 230-243, 254-267`: `level` is the suffix-array sampling level
            pub fn new<T: AsRef<[C]>>(text: &Text<C, T>, level: usize) -> Result<Self, Error> {
                Ok($name(Core::build(text, <$flavor as Flavor>::KIND, level as c_int)?, PhantomData))
            }
        }
    };
}
macro_rules! multi_pieces {
    ($name:ident, $search:ident) => {
        impl<C: Character> SearchIndexWithMultiPieces<C> for $name<C> {
            fn search_prefix<K>(&self, pattern: K) -> impl Search<'_, C>
            where
                K: AsRef<[C]>,
            {
                $name::search_prefix(self, pattern)
            }
            fn search_suffix<K>(&self, pattern: K) -> impl Search<'_, C>
            where
                K: AsRef<[C]>,
            {
                $name::search_suffix(self, pattern)
            }
            fn search_exact<K>(&self, pattern: K) -> impl Search<'_, C>
            where
                K: AsRef<[C]>,
            {
                $name::search_exact(self, pattern)
            }
        }
        impl<C: Character> $name<C> {
            /// matches at the start of a piece
            pub fn search_prefix<K: AsRef<[C]>>(&self, pattern: K) -> $search<'_, C> {
                SearchOf::fresh(&self.0, ffi::FMX_SEARCH_PREFIX, pattern)
            }
            /// matches at the end of a piece
            pub fn search_suffix<K: AsRef<[C]>>(&self, pattern: K) -> $search<'_, C> {
                SearchOf::fresh(&self.0, ffi::FMX_SEARCH_SUFFIX, pattern)
            }
            /// matches that are a whole piece
            pub fn search_exact<K: AsRef<[C]>>(&self, pattern: K) -> $search<'_, C> {
                SearchOf::fresh(&self.0, ffi::FMX_SEARCH_EXACT, pattern)
            }
            /// `HasMultiPieces::pieces_count` (src/multi_pieces.rs:220-222)
            pub fn pieces_count(&self) -> usize {
                unsafe { ffi::fmx_index_pieces_count(self.0.h) as usize }
            }
        }
    };
}

index_type!(FMIndex, FmCount, FMIndexSearch, FMIndexMatch, "FMIndex, count only (`src/frontend.rs:110`).");
index_type!(FMIndexWithLocate, FmLocate, FMIndexSearchWithLocate, FMIndexMatchWithLocate, "FMIndex with locate support (`src/frontend.rs:124-126`).");
index_type!(RLFMIndex, RlCount, RLFMIndexSearch, RLFMIndexMatch, "RLFMIndex, count only (`src/frontend.rs:139`).");
index_type!(RLFMIndexWithLocate, RlLocate, RLFMIndexSearchWithLocate, RLFMIndexMatchWithLocate, "RLFMIndex with locate support (`src/frontend.rs:153-155`).");
index_type!(FMIndexMultiPieces, MultiCount, FMIndexMultiPiecesSearch, FMIndexMultiPiecesMatch, "Multi-piece index, count only (`src/frontend.rs:168-170`).");
index_type!(FMIndexMultiPiecesWithLocate, MultiLocate, FMIndexMultiPiecesSearchWithLocate, FMIndexMultiPiecesMatchWithLocate,
            "Multi-piece index with locate support (`src/frontend.rs:184-186`).");
count_only_new!(FMIndex, FmCount);
count_only_new!(RLFMIndex, RlCount);
count_only_new!(FMIndexMultiPieces, MultiCount);
with_locate_new!(FMIndexWithLocate, FmLocate);
with_locate_new!(RLFMIndexWithLocate, RlLocate);
with_locate_new!(FMIndexMultiPiecesWithLocate, MultiLocate);
multi_pieces!(FMIndexMultiPieces, FMIndexMultiPiecesSearch);
multi_pieces!(FMIndexMultiPiecesWithLocate, FMIndexMultiPiecesSearchWithLocate);

#[cfg(test)]
mod tests {
    //! the reference's own API tests (tests/test_api.rs, tests/test_fmindex.rs:6-24, README.md:49-64), verbatim in spirit
    use super::*;

    fn len<T: SearchIndex<u8>>(index: &T) -> usize {
        index.len()
    }
    fn size<T: SearchIndex<u8>>(index: &T) -> usize {
        index.heap_size()
    }
    fn count_generic<T: SearchIndex<u8>>(index: &T, p: &str) -> usize {
        index.search(p).count()
    }

    #[test]
    fn traits_are_usable_generically() {
        let text = Text::new("text\0".as_bytes());
        let a = FMIndexWithLocate::new(&text, 2).unwrap();
        let b = FMIndex::new(&text).unwrap();
        let c = RLFMIndexWithLocate::new(&text, 2).unwrap();
        assert_eq!((len(&a), len(&b), len(&c)), (5, 5, 5));
        assert!(size(&a) > 0 && size(&b) > 0 && size(&c) > 0);
        assert_eq!(count_generic(&a, "t"), 2);
    }

    #[test]
    fn small_locate() {
        let index = FMIndexWithLocate::new(&Text::new("a\0".as_bytes()), 2).unwrap();
        let search = index.search("a");
        assert_eq!(search.count(), 1);
        assert_eq!(search.iter_matches().map(|m| m.locate()).collect::<Vec<_>>(), vec![0]);
    }

    #[test]
    fn invalid_text() {
        match FMIndex::new(&Text::new("\0a\0".as_bytes())) {
            Err(Error::InvalidText(msg)) => assert!(msg.contains("must not start")),
            _ => panic!("expected InvalidText"),
        }
    }
}
