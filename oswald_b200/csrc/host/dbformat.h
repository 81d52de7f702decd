/* dbformat.h - the GPU database layout: length-binned chunk streams.
 *
 * Replaces the reference's 32-lane interleaved groups (assemble_db_chunks, reference
 * sequences.c:1035-1088) and its measured-GCUPS FPGA/host split (sequences.c:842-863).
 *
 * The canonical database (stable ascending length order, sequences.c:1130-1225) is cut into
 * CHUNKS of consecutive whole sequences holding about `chunk_cols` residues each (one longer
 * sequence is a chunk of its own; the chunks of the shortest sequences, which the GPU takes last,
 * are a quarter and then a sixteenth of that size to even out the end of a launch).  Because the order is by length, a chunk is a length bin:
 * all its sequences have (nearly) the same length.  A chunk is stored as a COLUMN STREAM, one
 * byte per residue:
 *      bits 0-4  residue code 0..23
 *      bit  5    OSW_COL_FIRST  first column of a sequence (DP state restarts here)
 *      bit  6    OSW_COL_LAST   last column of a sequence (its score is complete here)
 * Empty sequences form a chunk of their own without columns (their scores stay 0).
 * Chunk streams start on 128-byte boundaries (padding bytes are OSW_COL_PADBYTE), so a warp reads a
 * stream with coalesced 128-bit loads.
 *
 * The shard also exists as a PAIR STREAM for searches with a single query (or very unequal
 * query sets), where the two 16-bit halves of the kernel's words score two DIFFERENT database
 * sequences against the same query rows: the shard's sequences 2p and 2p+1 (neighbours in
 * length) are zipped column by column, the shorter one padded with OSW_COL_PADBYTE (which scores
 * 0 against everything and so cannot raise a maximum).  Two bytes per column:
 *      byte 0   residue of the first sequence | FIRST / LAST flags of the pair's columns
 *      byte 1   residue of the second sequence (padding for the single last sequence)
 * The pairs are grouped into PAIR CHUNKS with a directory of their own (about chunk_cols / 2 pair
 * columns each: the same work as a plain chunk); pair chunks start on 64-column (128-byte)
 * boundaries.  Chunk c of the database is dealt to shard
 * (c mod n_shards): every shard gets the same mix of lengths and the same number of
 * residues to within one chunk - the residue-balanced split across GPUs.
 */
#ifndef OSW_DBFORMAT_H
#define OSW_DBFORMAT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OSW_COL_CODE  0x1f
#define OSW_COL_FIRST 0x20
#define OSW_COL_LAST  0x40
#define OSW_COL_PADBYTE 23           /* pad residue, no flags */
#define OSW_CHUNK_ALIGN 128
#define OSW_CHUNK_COLS_DEFAULT 8192

typedef struct osw_chunk {
    uint64_t stream_off;   /* byte offset of the first column in the shard stream */
    uint32_t n_cols;       /* residues in the chunk */
    uint32_t n_seqs;       /* whole sequences in the chunk */
    uint32_t seq0;         /* shard-local index of its first sequence */
    uint32_t canon0;       /* canonical index of its first sequence (the rest follow) */
    uint64_t pair_off;     /* pair directory: column offset of the chunk in the PAIR stream (2 bytes per column) */
    uint32_t n_pair_cols;  /* pair directory: columns of the chunk in the pair stream */
    uint32_t reserved;
} osw_chunk;

typedef struct osw_shard {
    uint64_t   n_seqs;         /* sequences in this shard */
    uint64_t   n_residues;     /* residues in this shard */
    uint64_t   stream_bytes;   /* size of stream (multiple of OSW_CHUNK_ALIGN) */
    uint32_t   n_chunks;
    uint32_t   max_len;        /* longest sequence in the shard */
    int        external_streams; /* stream / pair_stream came from a caller's allocator: not freed here */
    uint8_t   *stream;         /* column stream, chunks back to back */
    uint64_t   pair_cols;      /* columns of the pair stream (multiple of 64) */
    uint8_t   *pair_stream;    /* 2 * pair_cols bytes */
    osw_chunk *chunks;         /* in DESCENDING length order (longest work is handed out first) */
    osw_chunk *pair_chunks;    /* the pair directory (same order; stream_off / n_cols unused) */
    uint32_t   n_pair_chunks;
    uint32_t  *canon;          /* canon[local] = canonical index */
    uint64_t  *seq_off;        /* seq_off[local] = stream offset of the sequence's first column */
    uint32_t  *seq_len;        /* seq_len[local] */
} osw_shard;

/* Number of chunks the whole canonical database is cut into. */
uint64_t osw_count_chunks(const uint64_t *offsets, uint64_t n_seqs, uint32_t chunk_cols);

/* Build shard `shard` of `n_shards` (chunk c belongs to shard c mod n_shards).
 * Returns 0, -1 on allocation failure / bad arguments, -2 if a residue code is not in 0..23. */
int osw_shard_build(const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                    uint32_t shard, uint32_t n_shards, uint32_t chunk_cols, osw_shard *out);
/* Same, with the two stream buffers taken from `alloc(bytes, user)` (e.g. pinned host memory);
 * such buffers are not freed by osw_shard_free. */
typedef void *(*osw_alloc_fn)(size_t bytes, void *user);
int osw_shard_build_ex(const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                       uint32_t shard, uint32_t n_shards, uint32_t chunk_cols, int with_pair,
                       osw_alloc_fn alloc, void *alloc_user, osw_shard *out);
/* ---- the same layout on disk: X.osw ----------------------------------------------------------
 * `-O preprocess` writes it next to the reference's triple (X.info / X.seq / X.desc, sequences.c:177-208);
 * a search then maps it and copies chunk streams straight into pinned memory instead of re-laying
 * the database out (flags, padding, directories) at every start.  One file serves any number of
 * GPUs: it holds ALL chunks of the canonical database in order, and shard s of n takes every n-th.
 * Little-endian, versioned; a reader rejects other versions, and the caller falls back to X.seq.
 *   header   128 bytes (osw_dbfile_header)
 *   chunks   n_chunks x osw_chunk, DESCENDING length order, offsets relative to the stream section
 *   lengths  n_seqs x uint32, canonical order
 *   stream   stream_bytes: chunk column streams back to back, each padded to OSW_CHUNK_ALIGN  */
#define OSW_DBFILE_MAGIC "OSWB200"         /* 7 characters + NUL */
#define OSW_DBFILE_VERSION 1u
typedef struct osw_dbfile_header {
    char     magic[8];
    uint32_t version;
    uint32_t chunk_cols;       /* the work-unit size the chunks were cut with */
    uint64_t n_seqs, n_residues, n_chunks, stream_bytes;
    uint32_t max_len, chunk_align;
    uint64_t off_chunks, off_lengths, off_stream, file_bytes;
    uint64_t checksum;         /* FNV-1a of the chunk directory and the lengths */
    uint8_t  reserved[32];
} osw_dbfile_header;
typedef struct osw_dbfile {
    osw_dbfile_header h;
    const osw_chunk *chunks;
    const uint32_t  *lengths;
    const uint8_t   *stream;
    void *map; size_t map_size;
} osw_dbfile;
/* 0, -1 cannot create / write, -2 bad residue code, -3 allocation failure */
int  osw_dbfile_write(const char *path, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs, uint32_t chunk_cols);
/* 0, -1 cannot open, -2 not an X.osw file of this version, -3 truncated or corrupt */
int  osw_dbfile_open(const char *path, osw_dbfile *f);
void osw_dbfile_close(osw_dbfile *f);
/* Shard `shard` of `n_shards` from a mapped file: the same osw_shard osw_shard_build_ex makes from the
 * canonical arrays with the file's chunk_cols (tested).  0 or -1. */
int  osw_shard_from_file(const osw_dbfile *f, uint32_t shard, uint32_t n_shards, osw_alloc_fn alloc, void *alloc_user, osw_shard *out);

/* Fills pair[2 * s->pair_cols] from the shard's plain stream (for a deferred pair stream). */
void osw_shard_fill_pair(const osw_shard *s, const uint8_t *stream, uint8_t *pair);
void osw_shard_free(osw_shard *s);

#ifdef __cplusplus
}
#endif
#endif
