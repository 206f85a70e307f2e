//! UNVERIFIED SOURCE (no Rust toolchain in the build image).
//!
//! Safe wrappers that keep the crate's API surface for bulk use:
//! `GpuEncoder<P>::{encode, decode, rev_comp, get, get_prefix}` are `Encoding<P, B>` / `Kmer<P, K, B>` (src/encoding/mod.rs:14-23,
//! src/kmer.rs:12-53) for any word type `P` over slices of
//! k-mers; `ReadBatch::canonical_kmers` is `CanonicalKmerIterator` (src/naive_impl/canonical_kmer_iterator.rs)
//! + `get_canonical_word` + `hash_one(&LexHasherState::new(k), ..)` over a whole batch of reads.
use kmers_b200_sys as sys;
use std::{ffi::CStr, marker::PhantomData, ptr};

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub msg: String,
}

fn check(ctx: *const sys::kmb_ctx, rc: i32) -> Result<(), Error> {
    if rc == sys::KMB_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::kmb_last_error(ctx)) }.to_string_lossy().into_owned();
    if rc == sys::KMB_ERR_PANIC {
        panic!("{msg}"); // same contract as the reference: these arguments panic
    }
    Err(Error { code: rc, msg })
}

pub struct Context {
    raw: *mut sys::kmb_ctx,
}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        check(ptr::null(), unsafe { sys::kmb_ctx_create(device, ptr::null_mut(), &mut raw) })?;
        Ok(Self { raw })
    }

    /// Upload fixed-length reads (pinned staging -> device).
    pub fn upload<'c>(&'c mut self, bases: &[u8], fixed_len: usize) -> Result<ReadBatch<'c>, Error> {
        let n_reads = bases.len() / fixed_len;
        check(self.raw, unsafe {
            sys::kmb_batch_upload(self.raw, bases.as_ptr(), bases.len() as u64, ptr::null(), n_reads as u64, fixed_len as u64)
        })?;
        Ok(ReadBatch { ctx: self.raw, _life: PhantomData })
    }

    /// Upload ragged reads (CSR offsets).
    pub fn upload_ragged<'c>(&'c mut self, bases: &[u8], offsets: &[u64]) -> Result<ReadBatch<'c>, Error> {
        check(self.raw, unsafe {
            sys::kmb_batch_upload(self.raw, bases.as_ptr(), bases.len() as u64, offsets.as_ptr(), (offsets.len() - 1) as u64, 0)
        })?;
        Ok(ReadBatch { ctx: self.raw, _life: PhantomData })
    }

    /// `Kmer::get_reverse_complement_word` on every word (src/naive_impl/kmer.rs:138-147).
    pub fn reverse_complement_words(&mut self, words: &[u64], k: u8) -> Result<Vec<u64>, Error> {
        let mut out = vec![0u64; words.len()];
        check(self.raw, unsafe { sys::kmb_reverse_complement_words(self.raw, k as u32, words.as_ptr(), out.as_mut_ptr(), words.len() as u64) })?;
        Ok(out)
    }

    /// `hash_one(&LexHasherState::new(k), kmer)` on every word (src/naive_impl/hash.rs:10-20).
    pub fn lex_hash_words(&mut self, words: &[u64], k: u8) -> Result<Vec<u64>, Error> {
        let mut out = vec![0u64; words.len()];
        check(self.raw, unsafe { sys::kmb_lexhash_words(self.raw, k as u32, words.as_ptr(), out.as_mut_ptr(), words.len() as u64) })?;
        Ok(out)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::kmb_ctx_destroy(self.raw) };
    }
}

/// Dense-slot result: slot `r * (L - k + 1) + pos`; `u64::MAX` where the iterator skips the window.
pub struct CanonicalKmers {
    pub canon: Vec<u64>,
    pub hash: Vec<u64>,
    pub digest: sys::kmb_digest,
}

pub struct ReadBatch<'c> {
    ctx: *mut sys::kmb_ctx,
    _life: PhantomData<&'c mut Context>,
}

impl<'c> ReadBatch<'c> {
    pub fn canonical_kmers(&self, k: u8) -> Result<CanonicalKmers, Error> {
        let mut n = 0u64;
        check(self.ctx, unsafe { sys::kmb_batch_num_slots(self.ctx, k as u32, &mut n) })?;
        let mut canon = vec![0u64; n as usize];
        let mut hash = vec![0u64; n as usize];
        let mut digest = sys::kmb_digest::default();
        check(self.ctx, unsafe {
            sys::kmb_extract_canonical(self.ctx, k as u32, 0, canon.as_mut_ptr(), hash.as_mut_ptr(), ptr::null_mut(), ptr::null_mut(), &mut digest)
        })?;
        Ok(CanonicalKmers { canon, hash, digest })
    }
}

/// Word types of the generic `Kmer<P, K, B>` (src/utils.rs:4-24: u8 .. u128), as the library sees them: plain
/// little-endian byte images of `BITS` bits.
pub trait Word: Copy + Default {
    const BITS: u32;
}
macro_rules! word { ($($t:ty),*) => { $(impl Word for $t { const BITS: u32 = (std::mem::size_of::<$t>() * 8) as u32; })* } }
word!(u8, u16, u32, u64, u128);

/// Batched `Encoding<P, B>` (src/encoding/mod.rs:14-23): the `Naive` discriminant byte (src/encoding/naive.rs:49-74) or
/// Xor10 selects the code; `P` is any of the reference's word types.
pub struct GpuEncoder<'c, P: Word> {
    pub ctx: &'c mut Context,
    pub enc: i32,
    _word: PhantomData<P>,
}

impl<'c, P: Word> GpuEncoder<'c, P> {
    pub fn naive(ctx: &'c mut Context, enc: kmers::encoding::Naive) -> Self {
        Self { ctx, enc: enc as u8 as i32, _word: PhantomData }
    }
    pub fn xor10(ctx: &'c mut Context) -> Self {
        Self { ctx, enc: sys::KMB_ENC_XOR10, _word: PhantomData }
    }

    /// `Encoding::encode` for `seqs.len() / K` k-mers of `K` ASCII bytes each.  The library writes
    /// `word_for_k::<P, K>()` words per k-mer back to back (src/kmer.rs:67-69), so `B` must be exactly that:
    /// a larger `B` would silently misalign the rows.
    pub fn encode<const K: usize, const B: usize>(&mut self, seqs: &[u8]) -> Result<Vec<[P; B]>, Error> {
        let bases_per_word = (P::BITS / 2) as usize;
        assert_eq!(B, (K + bases_per_word - 1) / bases_per_word, "B must equal word_for_k::<P, K>()");
        let n = seqs.len() / K;
        let mut out = vec![[P::default(); B]; n];
        let raw = self.ctx.raw;
        check(raw, unsafe { sys::kmb_batch_upload(raw, seqs.as_ptr(), seqs.len() as u64, ptr::null(), n as u64, K as u64) })?;
        check(raw, unsafe { sys::kmb_pack(raw, self.enc, P::BITS, out.as_mut_ptr() as *mut _, ptr::null_mut()) })?;
        Ok(out)
    }

    /// `Encoding::rev_comp::<K>` on every array.
    pub fn rev_comp<const K: usize, const B: usize>(&mut self, arrays: &[[P; B]]) -> Result<Vec<[P; B]>, Error> {
        let mut out = vec![[P::default(); B]; arrays.len()];
        let raw = self.ctx.raw;
        check(raw, unsafe {
            sys::kmb_revcomp_words(raw, self.enc, K as u32, P::BITS, B as u32, arrays.as_ptr() as *const _, out.as_mut_ptr() as *mut _, arrays.len() as u64)
        })?;
        Ok(out)
    }

    /// `Encoding::decode` of every array (all positions, padding included, like the reference).
    pub fn decode<const B: usize>(&mut self, arrays: &[[P; B]]) -> Result<Vec<Vec<u8>>, Error> {
        let per = B * (P::BITS / 2) as usize;
        let mut flat = vec![0u8; arrays.len() * per];
        let raw = self.ctx.raw;
        check(raw, unsafe {
            sys::kmb_unpack(raw, self.enc, P::BITS, arrays.as_ptr() as *const _, arrays.len() as u64, B as u32, per as u32, flat.as_mut_ptr())
        })?;
        Ok(flat.chunks(per).map(|c| c.to_vec()).collect())
    }

    /// `Kmer::<P, K, B>::get(index)` on every array (src/kmer.rs:46-48).
    pub fn get<const B: usize>(&mut self, arrays: &[[P; B]], index: usize) -> Result<Vec<u8>, Error> {
        let mut out = vec![0u8; arrays.len()];
        let raw = self.ctx.raw;
        check(raw, unsafe { sys::kmb_kmer_get(raw, P::BITS, B as u32, arrays.as_ptr() as *const _, arrays.len() as u64, index as u32, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `Kmer::<P, K, B>::get_prefix(len)` on every array (src/kmer.rs:50-52; 2 len + 1 bits, as the reference).
    pub fn get_prefix<const B: usize>(&mut self, arrays: &[[P; B]], len: usize) -> Result<Vec<P>, Error> {
        let mut out = vec![P::default(); arrays.len()];
        let raw = self.ctx.raw;
        check(raw, unsafe {
            sys::kmb_kmer_get_prefix(raw, P::BITS, B as u32, arrays.as_ptr() as *const _, arrays.len() as u64, len as u32, out.as_mut_ptr() as *mut _)
        })?;
        Ok(out)
    }
}

impl Context {
    /// Reads in host memory -> canonical words + LexHashes in host vectors, through the pipelined path
    /// (host packing, H2D, kernel and D2H overlapped).
    pub fn canonical_kmers_host(&mut self, bases: &[u8], fixed_len: usize, k: u8) -> Result<CanonicalKmers, Error> {
        let n_reads = bases.len() / fixed_len;
        let n = n_reads * (fixed_len + 1).saturating_sub(k as usize);
        let (mut canon, mut hash) = (vec![0u64; n], vec![0u64; n]);
        let mut digest = sys::kmb_digest::default();
        check(self.raw, unsafe {
            sys::kmb_extract_canonical_host(self.raw, bases.as_ptr(), n_reads as u64, fixed_len as u64, k as u32, 0, canon.as_mut_ptr(), hash.as_mut_ptr(), &mut digest)
        })?;
        Ok(CanonicalKmers { canon, hash, digest })
    }

    /// `Kmer::append_base` on every word (src/naive_impl/kmer.rs:97-102): (shifted words, bases shifted off).
    pub fn append_base_words(&mut self, words: &[u64], bases: &[u8], k: u8) -> Result<(Vec<u64>, Vec<u8>), Error> {
        assert_eq!(words.len(), bases.len());
        let (mut out, mut dropped) = (vec![0u64; words.len()], vec![0u8; words.len()]);
        check(self.raw, unsafe {
            sys::kmb_append_base_words(self.raw, k as u32, words.as_ptr(), bases.as_ptr(), 0, out.as_mut_ptr(), dropped.as_mut_ptr(), words.len() as u64)
        })?;
        Ok((out, dropped))
    }

    /// `Kmer::sub_kmer_word` on every word (src/naive_impl/kmer.rs:150-161).
    pub fn sub_kmer_words(&mut self, words: &[u64], k: u8, pos: usize, width: usize) -> Result<Vec<u64>, Error> {
        let mut out = vec![0u64; words.len()];
        check(self.raw, unsafe { sys::kmb_sub_kmer_words(self.raw, k as u32, pos as u32, width as u32, words.as_ptr(), out.as_mut_ptr(), words.len() as u64) })?;
        Ok(out)
    }
}
