//! `extern "C"` declarations for include/fmx.h (one per symbol the facade uses).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct fmx_index {
    _private: [u8; 0],
}

/// several GPUs of one process behind one handle (include/fmx.h, "multi-GPU")
#[repr(C)]
pub struct fmx_group {
    _private: [u8; 0],
}
pub const FMX_GROUP_REPLICATE: c_int = 0;
pub const FMX_GROUP_BY_PIECE: c_int = 1;

pub const FMX_KIND_FM: c_int = 0;
pub const FMX_KIND_RLFM: c_int = 1;
pub const FMX_KIND_MULTI: c_int = 2;
pub const FMX_SEARCH: c_int = 0;
pub const FMX_SEARCH_PREFIX: c_int = 1;
pub const FMX_SEARCH_SUFFIX: c_int = 2;
pub const FMX_SEARCH_EXACT: c_int = 3;
pub const FMX_LEVEL_COUNT_ONLY: c_int = -1;
pub const FMX_ERR_INVALID_TEXT: c_int = -1;
pub const FMX_ERR_PATTERN_CHAR: c_int = -5;
pub const FMX_ERR_CAPACITY: c_int = -9;
pub const FMX_MODE_AUTO: c_int = 0;
pub const FMX_MODE_COMPACT: c_int = 1;
pub const FMX_MODE_RICH: c_int = 2;

/// struct fmx_query (include/fmx.h): one descriptor for the fused count + locate call
#[repr(C)]
pub struct fmx_query {
    pub mode: c_int,
    pub packed_bits: u32,
    pub patterns: *const c_void,
    pub pat_off: *const u64,
    pub fixed_len: u64,
    pub npat: u64,
    pub out_width: u32,
    pub reserved: u32,
    pub out_s: *mut u64,
    pub out_e: *mut u64,
    pub counts: *mut c_void,
    pub hit_off: *mut c_void,
    pub positions: *mut c_void,
    pub piece_ids: *mut c_void,
    pub capacity: u64,
}

extern "C" {
    pub fn fmx_last_error() -> *const c_char;
    pub fn fmx_free(p: *mut c_void);
    pub fn fmx_index_build(text: *const c_void, n: u64, char_width: u32, max_character: u64, kind: c_int,
                           level: c_int, device: c_int, out: *mut *mut fmx_index) -> c_int;
    pub fn fmx_index_build_ex(text: *const c_void, n: u64, char_width: u32, max_character: u64, kind: c_int,
                              level: c_int, device: c_int, mode: c_int, out: *mut *mut fmx_index) -> c_int;
    pub fn fmx_index_free(idx: *mut fmx_index);
    pub fn fmx_query_batch(idx: *const fmx_index, q: *const fmx_query, total_hits: *mut u64) -> c_int;
    pub fn fmx_index_len(idx: *const fmx_index) -> u64;
    pub fn fmx_index_device_bytes(idx: *const fmx_index) -> u64;
    pub fn fmx_index_pieces_count(idx: *const fmx_index) -> u64;
    pub fn fmx_search_batch(idx: *const fmx_index, mode: c_int, pat: *const u8, pat_off: *const u64, fixed_len: u64,
                            npat: u64, init_s: *const u64, init_e: *const u64, out_s: *mut u64, out_e: *mut u64) -> c_int;
    pub fn fmx_locate_batch(idx: *const fmx_index, prefix_only: c_int, s: *const u64, e: *const u64, npat: u64,
                            hit_off: *mut u64, positions: *mut *mut u64, piece_ids: *mut *mut u64) -> c_int;
    // one page of the hit list: the bounded-memory form of the lazy iter_matches (wrapper.rs:137-139, 203-217)
    pub fn fmx_locate_page(idx: *const fmx_index, s: *const u64, e: *const u64, npat: u64, first_hit: u64, nhits: u64,
                           positions: *mut u64, piece_ids: *mut u64, total_hits: *mut u64) -> c_int;
    pub fn fmx_search_locate_batch(idx: *const fmx_index, mode: c_int, pat: *const u8, pat_off: *const u64,
                                   fixed_len: u64, npat: u64, out_s: *mut u64, out_e: *mut u64, hit_off: *mut u64,
                                   positions: *mut u64, piece_ids: *mut u64, capacity: u64, total_hits: *mut u64) -> c_int;
    pub fn fmx_extract_batch(idx: *const fmx_index, rows: *const u64, nrows: u64, k: u32, forward: c_int,
                             out: *mut u8, out_len: *mut u32) -> c_int;
    pub fn fmx_rows_op(idx: *const fmx_index, op: c_int, rows: *const u64, nrows: u64, out: *mut u64) -> c_int;
    pub fn fmx_group_create(device_ids: *const c_int, ndev: c_int, group_mode: c_int, text: *const c_void, n: u64,
                            char_width: u32, max_character: u64, kind: c_int, level: c_int, index_mode: c_int,
                            out: *mut *mut fmx_group) -> c_int;
    pub fn fmx_group_free(g: *mut fmx_group);
    pub fn fmx_group_size(g: *const fmx_group) -> c_int;
    pub fn fmx_group_len(g: *const fmx_group) -> u64;
    pub fn fmx_group_pieces_count(g: *const fmx_group) -> u64;
    pub fn fmx_group_query_batch(g: *const fmx_group, q: *const fmx_query, total_hits: *mut u64) -> c_int;
}
