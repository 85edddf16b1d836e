/* h5lite.c -- see h5lite.h.  Format reference: the HDF5 File Format Specification v1.1 (superblock v0). */
#include "h5lite.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct h5lite_entry { const char *name; uint64_t ohdr; };
struct h5lite_file {
  unsigned char *buf;
  size_t size;
  uint64_t base;
  struct h5lite_entry *ent;
  size_t nent, cap;
};

static _Thread_local char g_err[256];
const char *h5lite_error(void) { return g_err; }
static int fail(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return -1; }

static uint64_t rd(const unsigned char *p, int n) {  /* little-endian unsigned of n bytes */
  uint64_t v = 0;
  for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}
static int in_range(const h5lite_file *f, uint64_t off, uint64_t len) { return off <= f->size && len <= f->size - off; }

static int add_entry(h5lite_file *f, const char *name, uint64_t ohdr) {
  if (f->nent == f->cap) {
    size_t cap = f->cap ? 2 * f->cap : 64;
    struct h5lite_entry *e = realloc(f->ent, cap * sizeof *e);
    if (!e) return fail("out of memory");
    f->ent = e; f->cap = cap;
  }
  f->ent[f->nent].name = name; f->ent[f->nent].ohdr = ohdr; f->nent++;
  return 0;
}

/* group B-tree (node type 0) -> symbol-table nodes -> (name offset in the local heap, object header) */
static int walk_group(h5lite_file *f, uint64_t node, uint64_t heap_data, int depth) {
  if (depth > 16 || !in_range(f, node, 24)) return fail("corrupt group B-tree");
  const unsigned char *p = f->buf + node;
  if (!memcmp(p, "TREE", 4)) {
    if (p[4] != 0) return fail("not a group B-tree");
    unsigned nent = (unsigned)rd(p + 6, 2);
    if (!in_range(f, node + 24, 8 + 16ull * nent)) return fail("corrupt group B-tree");
    for (unsigned i = 0; i < nent; ++i)
      if (walk_group(f, f->base + rd(p + 24 + 8 + 16 * i, 8), heap_data, depth + 1)) return -1;
    return 0;
  }
  if (!memcmp(p, "SNOD", 4)) {
    unsigned nsym = (unsigned)rd(p + 6, 2);
    if (!in_range(f, node + 8, 40ull * nsym)) return fail("corrupt symbol node");
    for (unsigned i = 0; i < nsym; ++i) {
      uint64_t noff = rd(p + 8 + 40 * i, 8), ohdr = rd(p + 16 + 40 * i, 8);
      if (!in_range(f, heap_data + noff, 1)) return fail("name outside the local heap");
      if (!memchr(f->buf + heap_data + noff, 0, f->size - (heap_data + noff))) return fail("unterminated name");
      if (add_entry(f, (const char *)f->buf + heap_data + noff, f->base + ohdr)) return -1;
    }
    return 0;
  }
  return fail("unknown group node signature");
}

/* iterate the messages of a version-1 object header (following continuation blocks) */
typedef int (*msg_fn)(void *ctx, unsigned type, const unsigned char *body, unsigned size);
static int for_messages(const h5lite_file *f, uint64_t addr, msg_fn fn, void *ctx) {
  if (!in_range(f, addr, 16)) return fail("object header outside the file");
  const unsigned char *h = f->buf + addr;
  if (h[0] != 1) return fail("unsupported object header version");
  unsigned nmsg = (unsigned)rd(h + 2, 2), seen = 0;
  uint64_t blk[32][2];
  int nblk = 1, cur = 0;
  blk[0][0] = addr + 16; blk[0][1] = rd(h + 8, 4);
  while (cur < nblk && seen < nmsg) {
    uint64_t pos = blk[cur][0], end = pos + blk[cur][1];
    ++cur;
    if (!in_range(f, pos, end - pos)) return fail("object header block outside the file");
    while (pos + 8 <= end && seen < nmsg) {
      unsigned type = (unsigned)rd(f->buf + pos, 2), size = (unsigned)rd(f->buf + pos + 2, 2);
      const unsigned char *body = f->buf + pos + 8;
      if (pos + 8 + size > end) return fail("message overruns its block");
      pos += 8 + size; ++seen;
      if (type == 0x0010) {               /* continuation */
        if (nblk < 32) { blk[nblk][0] = f->base + rd(body, 8); blk[nblk][1] = rd(body + 8, 8); ++nblk; }
      } else {
        int r = fn(ctx, type, body, size);
        if (r) return r;
      }
    }
  }
  return 0;
}

struct symtab { uint64_t btree, heap; int found; };
static int find_symtab(void *ctx, unsigned type, const unsigned char *body, unsigned size) {
  struct symtab *s = ctx;
  if (type == 0x0011 && size >= 16) { s->btree = rd(body, 8); s->heap = rd(body + 8, 8); s->found = 1; }
  return 0;
}

h5lite_file *h5lite_open(const char *path) {
  static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  FILE *fp = fopen(path, "rb");
  if (!fp) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return NULL; }
  h5lite_file *f = calloc(1, sizeof *f);
  if (!f) { fclose(fp); fail("out of memory"); return NULL; }
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  f->buf = n > 0 ? malloc((size_t)n) : NULL;
  if (!f->buf || fread(f->buf, 1, (size_t)n, fp) != (size_t)n) {
    fclose(fp); snprintf(g_err, sizeof g_err, "cannot read %s", path); h5lite_close(f); return NULL;
  }
  fclose(fp);
  f->size = (size_t)n;
  if (f->size < 96 || memcmp(f->buf, sig, 8)) { fail("not an HDF5 file (bad signature)"); h5lite_close(f); return NULL; }
  if (f->buf[8] != 0) { fail("unsupported superblock version (only v0)"); h5lite_close(f); return NULL; }
  if (f->buf[13] != 8 || f->buf[14] != 8) { fail("only 8-byte offsets/lengths supported"); h5lite_close(f); return NULL; }
  f->base = rd(f->buf + 24, 8);
  uint64_t root_ohdr = f->base + rd(f->buf + 56 + 8, 8);   /* root symbol-table entry at byte 56 */
  struct symtab st = {0, 0, 0};
  if (for_messages(f, root_ohdr, find_symtab, &st)) { h5lite_close(f); return NULL; }
  if (!st.found) { fail("root group has no symbol table message"); h5lite_close(f); return NULL; }
  uint64_t heap = f->base + st.heap;
  if (!in_range(f, heap, 32) || memcmp(f->buf + heap, "HEAP", 4)) { fail("bad local heap signature"); h5lite_close(f); return NULL; }
  uint64_t heap_data = f->base + rd(f->buf + heap + 24, 8);
  if (walk_group(f, f->base + st.btree, heap_data, 0)) { h5lite_close(f); return NULL; }
  return f;
}

void h5lite_close(h5lite_file *f) {
  if (!f) return;
  free(f->buf); free(f->ent); free(f);
}
size_t h5lite_count(const h5lite_file *f) { return f->nent; }
const char *h5lite_name(const h5lite_file *f, size_t i) { return i < f->nent ? f->ent[i].name : NULL; }

struct dsctx { const h5lite_file *f; h5lite_dataset d; int have_space, have_type, have_layout; size_t esize; uint64_t layout_bytes; };
static int dataset_msg(void *ctx, unsigned type, const unsigned char *b, unsigned size) {
  struct dsctx *c = ctx;
  if (type == 0x0001 && size >= 4) {                 /* dataspace v1/v2 */
    int ver = b[0], rank = b[1], off = ver == 1 ? 8 : 4;
    if (rank > 8 || (unsigned)(off + 8 * rank) > size) return fail("unsupported dataspace");
    c->d.rank = rank; c->d.count = 1;
    for (int i = 0; i < rank; ++i) { c->d.dims[i] = (size_t)rd(b + off + 8 * i, 8); c->d.count *= c->d.dims[i]; }
    c->have_space = 1;
  } else if (type == 0x0003 && size >= 8) {          /* datatype: class 0 fixed point, 1 floating point */
    int cls = b[0] & 0x0f, bits0 = b[1];
    unsigned es = (unsigned)rd(b + 4, 4);
    if (bits0 & 1) return fail("big-endian datasets not supported");
    int sgn = bits0 & 0x08;
    if (cls == 1) c->d.type = es == 8 ? H5LITE_F64 : es == 4 ? H5LITE_F32 : 0;
    else if (cls == 0)
      c->d.type = es == 4 ? (sgn ? H5LITE_I32 : H5LITE_U32) : es == 8 ? (sgn ? H5LITE_I64 : H5LITE_U64)
                : es == 1 ? (sgn ? H5LITE_I8 : H5LITE_U8) : es == 2 ? (sgn ? H5LITE_I16 : H5LITE_U16) : 0;
    else c->d.type = 0;
    if (!c->d.type) return fail("unsupported datatype");
    c->esize = es; c->have_type = 1;
  } else if (type == 0x0008 && size >= 4) {          /* data layout v3: compact (0) or contiguous (1) */
    if (b[0] != 3) return fail("unsupported layout version");
    if (b[1] == 1 && size >= 18) {
      uint64_t a = rd(b + 2, 8), n = rd(b + 10, 8);
      if (a == UINT64_MAX) c->d.data = NULL;
      else { if (!in_range(c->f, c->f->base + a, n)) return fail("dataset outside the file"); c->d.data = c->f->buf + c->f->base + a; }
      c->layout_bytes = n;
    } else if (b[1] == 0) {
      /* compact: 2-byte size, then the data inside this message */
      const uint64_t n = rd(b + 2, 2);
      if (n + 4 > size) return fail("compact dataset larger than its layout message");
      c->d.data = b + 4;
      c->layout_bytes = n;
    } else return fail("chunked layout not supported");
    c->have_layout = 1;
  }
  return 0;
}

int h5lite_find(const h5lite_file *f, const char *name, h5lite_dataset *out) {
  if (name[0] == '/') ++name;   /* the reference addresses datasets both as "NAME" and "/NAME" */
  for (size_t i = 0; i < f->nent; ++i) {
    if (strcmp(f->ent[i].name, name)) continue;
    struct dsctx c;
    memset(&c, 0, sizeof c);
    c.f = f;
    if (for_messages(f, f->ent[i].ohdr, dataset_msg, &c)) return -1;
    if (!c.have_space || !c.have_type || !c.have_layout) return fail("object is not a simple dataset");
    /* the element count comes from the dataspace, the bytes that exist from the layout: a truncated or inconsistent
       file must fail here, not read past the data */
    if (c.d.data && (c.esize == 0 || c.d.count > c.layout_bytes / c.esize)) {
      snprintf(g_err, sizeof g_err, "dataset %s: dataspace (%zu elements of %zu bytes) exceeds its storage (%llu bytes)", name,
               c.d.count, c.esize, (unsigned long long)c.layout_bytes);
      return -1;
    }
    *out = c.d;
    return 0;
  }
  snprintf(g_err, sizeof g_err, "dataset %s not found", name);
  return -1;
}

static double elem_as_double(const h5lite_dataset *d, size_t i) {
  const unsigned char *p = d->data;
  switch (d->type) {
    case H5LITE_F64: { double v; memcpy(&v, p + 8 * i, 8); return v; }
    case H5LITE_F32: { float v; memcpy(&v, p + 4 * i, 4); return v; }
    case H5LITE_I32: { int32_t v; memcpy(&v, p + 4 * i, 4); return v; }
    case H5LITE_U32: { uint32_t v; memcpy(&v, p + 4 * i, 4); return v; }
    case H5LITE_I64: { int64_t v; memcpy(&v, p + 8 * i, 8); return (double)v; }
    case H5LITE_U64: { uint64_t v; memcpy(&v, p + 8 * i, 8); return (double)v; }
    case H5LITE_I16: { int16_t v; memcpy(&v, p + 2 * i, 2); return v; }
    case H5LITE_U16: { uint16_t v; memcpy(&v, p + 2 * i, 2); return v; }
    case H5LITE_I8: return (signed char)p[i];
    default: return p[i];
  }
}

int h5lite_read_double(const h5lite_file *f, const char *name, double *out, size_t count) {
  h5lite_dataset d;
  if (h5lite_find(f, name, &d)) return -1;
  if (d.count != count) { snprintf(g_err, sizeof g_err, "dataset %s holds %zu elements, %zu requested", name, d.count, count); return -1; }
  if (!d.data) { memset(out, 0, count * sizeof *out); return 0; }
  if (d.type == H5LITE_F64) { memcpy(out, d.data, count * sizeof *out); return 0; }
  for (size_t i = 0; i < count; ++i) out[i] = elem_as_double(&d, i);
  return 0;
}

int h5lite_read_int(const h5lite_file *f, const char *name, int *out, size_t count) {
  h5lite_dataset d;
  if (h5lite_find(f, name, &d)) return -1;
  if (d.count != count) { snprintf(g_err, sizeof g_err, "dataset %s holds %zu elements, %zu requested", name, d.count, count); return -1; }
  if (!d.data) { memset(out, 0, count * sizeof *out); return 0; }
  if (d.type == H5LITE_I32) { memcpy(out, d.data, count * sizeof *out); return 0; }
  for (size_t i = 0; i < count; ++i) out[i] = (int)elem_as_double(&d, i);
  return 0;
}
