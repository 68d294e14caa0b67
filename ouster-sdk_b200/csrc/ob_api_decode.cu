// ob_api_decode.cu -- C-ABI glue of the packet-decode path (ob_decoder_*, ob_decode_frames).
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "ob_api_common.h"
#include "ob_encode.h"

using namespace ob;

struct ob_decoder {
    int device;
    DecodeLayout L;
};

static bool valid_elem_size(uint32_t es) { return es == 1 || es == 2 || es == 4 || es == 8 || es == 6; }

static DecodeField to_dev(const ob_field_desc& f) {
    DecodeField d;
    d.offset = f.offset;
    d.elem_size = f.elem_size;
    d.mask = f.mask;
    d.shift = f.shift;
    d.range_return = f.range_return;
    d.zero_pattern = f.zero_pattern;
    d.pad = 0;
    return d;
}

// Frames of independent sensor streams carry their own LUTs.  The frame table is only a work list
// (every entry holds its own output pointers), so it may be reordered freely: keeping the frames
// of one LUT together lets each 6-12 MB table stay L2-resident while its frames are processed.
static void group_by_lut(std::vector<DecodeFrame>& frames) {
    bool any = false;
    for (const DecodeFrame& f : frames) any |= f.lut_dir != nullptr;
    if (!any) return;
    std::stable_sort(frames.begin(), frames.end(), [](const DecodeFrame& a, const DecodeFrame& b) {
        return reinterpret_cast<uintptr_t>(a.lut_dir) < reinterpret_cast<uintptr_t>(b.lut_dir);
    });
}

// TMA descriptors of a LUT for the pipelined K2 (null: that kernel is not applicable / not available)
static const void* maps_for(const DecodeLayout& L, int device, const void* dir, const void* off, int dtype) {
    uint32_t bw = 0, bh = 0;
    if (!dir || !off || !decode_pipe_box(L, device, dtype, &bw, &bh)) return nullptr;
    return lut_tensor_maps(dir, off, dtype, L.H, L.W, bw, bh, device);
}

extern "C" {

ob_status ob_decoder_create(const ob_packet_layout* layout, const ob_field_desc* fields,
                            size_t n_fields, int device, ob_decoder** out) {
    if (!layout || !out || (n_fields && !fields)) return fail(OB_INVALID_ARGUMENT, "null pointer");
    if (n_fields > OB_MAX_FIELDS) return fail(OB_INVALID_ARGUMENT, "too many fields");
    if (layout->columns_per_packet == 0)
        return fail(OB_INVALID_ARGUMENT, "unexpected columns_per_packet: 0");  // lidar_frame.cpp:1250-1252
    if (layout->pixels_per_column == 0)
        return fail(OB_INVALID_ARGUMENT, "unexpected pixels_per_column: 0");   // lidar_frame.cpp:1253-1255
    if (layout->columns_per_frame == 0) return fail(OB_INVALID_ARGUMENT, "unexpected frame dimensions");
    if (layout->columns_per_packet > 64)
        return fail(OB_INVALID_ARGUMENT, "columns_per_packet above 64 is not supported");
    if (layout->packet_size > 65535)
        return fail(OB_INVALID_ARGUMENT, "lidar_packet_size cannot exceed 65535");  // parsing.cpp:471-473
    const uint64_t need = static_cast<uint64_t>(layout->packet_header_size) +
                          static_cast<uint64_t>(layout->columns_per_packet) * layout->col_size;
    if (need > layout->packet_size ||
        static_cast<uint64_t>(layout->col_header_size) +
                static_cast<uint64_t>(layout->pixels_per_column) * layout->channel_data_size >
            layout->col_size)
        return fail(OB_INVALID_ARGUMENT, "inconsistent packet layout");
    for (size_t i = 0; i < n_fields; ++i) {
        if (!valid_elem_size(fields[i].elem_size))
            return fail(OB_INVALID_ARGUMENT, "Dest type too small for specified field");
        if (fields[i].offset >= layout->channel_data_size + 8u && layout->channel_data_size > 0)
            return fail(OB_INVALID_ARGUMENT, "field offset outside the channel data block");
        if (fields[i].range_return >= OB_MAX_RETURNS)
            return fail(OB_INVALID_ARGUMENT, "range_return must be < 2");
        if (fields[i].range_return >= 0 && fields[i].elem_size != 4)
            return fail(OB_INVALID_ARGUMENT, "range fields must decode to uint32");
    }
    ob_status rs = require_device(device);
    if (rs != OB_OK) return rs;
    ob_decoder* d = new ob_decoder;
    d->device = device;
    std::memset(&d->L, 0, sizeof(d->L));
    d->L.packet_header_size = layout->packet_header_size;
    d->L.col_header_size = layout->col_header_size;
    d->L.channel_data_size = layout->channel_data_size;
    d->L.col_size = layout->col_size;
    d->L.packet_size = layout->packet_size;
    d->L.cpp = layout->columns_per_packet;
    d->L.H = layout->pixels_per_column;
    d->L.W = layout->columns_per_frame;
    d->L.ts = to_dev(layout->col_timestamp);
    d->L.mid = to_dev(layout->col_measurement_id);
    d->L.status = to_dev(layout->col_status);
    d->L.n_fields = static_cast<uint32_t>(n_fields);
    for (size_t i = 0; i < n_fields; ++i) d->L.fields[i] = to_dev(fields[i]);
    *out = d;
    return OB_OK;
}

ob_status ob_decoder_destroy(ob_decoder* dec) {
    delete dec;
    return OB_OK;
}

ob_status ob_decode_frames(const ob_decoder* dec, const ob_decode_io* frames, size_t n_frames,
                           const ob_lut* lut, const int32_t* shifts, size_t n_shifts, ob_stream* s) {
    if (!dec || !s || (n_frames && !frames)) return fail(OB_INVALID_ARGUMENT, "null pointer");
    if (n_frames == 0) return OB_OK;
    const DecodeLayout& L = dec->L;
    const int device = stream_device(s);
    if (device != dec->device) return fail(OB_INVALID_ARGUMENT, "decoder and stream are on different devices");
    ob_status rs = require_device(device);
    if (rs != OB_OK) return rs;
    const void *ldir = nullptr, *loff = nullptr;
    int ldtype = OB_F32;
    if (lut) {
        size_t lh, lw;
        int ldev;
        lut_view(lut, &ldir, &loff, &ldtype, &lh, &lw, &ldev);
        if (lh != L.H || lw != L.W) return fail(OB_INVALID_ARGUMENT, "unexpected image dimensions");
        if (ldev != device) return fail(OB_INVALID_ARGUMENT, "lut and stream are on different devices");
    }
    std::vector<uint16_t> sh;
    if (shifts) {
        if (n_shifts != L.H) return fail(OB_INVALID_ARGUMENT, "image height does not match shifts size");
        if (L.H > static_cast<uint32_t>(kMaxRows))
            return fail(OB_INVALID_ARGUMENT, "fused destagger supports at most 512 rows");
        reduce_shifts(shifts, L.H, L.W, 0, sh);
    }
    const size_t n_px = static_cast<size_t>(L.H) * L.W;
    cudaStream_t st = stream_handle(s);
    Staging stg(st);
    std::vector<DecodeFrame> hf(n_frames);
    bool vec_ok = true, frame_maps_ok = true, any_xyz = false;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (lut && (!al16(ldir) || !al16(loff))) vec_ok = false;
    for (size_t i = 0; i < n_frames; ++i) {
        const ob_decode_io& io = frames[i];
        DecodeFrame& f = hf[i];
        std::memset(&f, 0, sizeof(f));
        if (io.n_slots > 0 && !io.packets) return fail(OB_INVALID_ARGUMENT, "null packet buffer");
        if (io.n_slots > 0 && io.packet_stride < L.packet_size)
            return fail(OB_INVALID_ARGUMENT, "packet_stride smaller than the lidar packet size");
        if (io.n_slots > (1u << 20)) return fail(OB_INVALID_ARGUMENT, "too many packet slots");
        const void* d = nullptr;
        cudaError_t e = cudaSuccess;
        if (io.n_slots > 0) {
            e = stg.in(io.packets, (io.n_slots - 1) * io.packet_stride + L.packet_size, &d);
            if (e != cudaSuccess) return fail_cuda(e, "stage packets");
        }
        f.packets = static_cast<const uint8_t*>(d);
        f.packet_stride = io.packet_stride;
        f.n_slots = static_cast<uint32_t>(io.n_slots);
        f.flags = 0;
        if (!io.col_src) f.flags |= 1u;
        if (d && al16(d) && io.packet_stride % 16 == 0 && L.packet_size % 16 == 0) f.flags |= 2u;
        if (io.col_src) {
            e = stg.in(io.col_src, static_cast<size_t>(L.W) * 4, &d);
            if (e != cudaSuccess) return fail_cuda(e, "stage column map");
            f.col_src = static_cast<const int32_t*>(d);
        }
        void* o = nullptr;
        for (uint32_t k = 0; k < L.n_fields; ++k) {
            if (!io.fields[k]) continue;
            e = stg.out(io.fields[k], n_px * L.fields[k].elem_size, &o);
            if (e != cudaSuccess) return fail_cuda(e, "stage field output");
            f.fields[k] = o;
        }
        {
            bool allf = L.n_fields > 0;
            for (uint32_t k = 0; k < L.n_fields; ++k) allf = allf && f.fields[k] != nullptr;
            if (allf) f.flags |= 4u;  // every decoder field has an output image
        }
        if (io.timestamp) {
            e = stg.out(io.timestamp, static_cast<size_t>(L.W) * 8, &o);
            if (e != cudaSuccess) return fail_cuda(e, "stage timestamp");
            f.timestamp = static_cast<uint64_t*>(o);
        }
        if (io.measurement_id) {
            e = stg.out(io.measurement_id, static_cast<size_t>(L.W) * 2, &o);
            if (e != cudaSuccess) return fail_cuda(e, "stage measurement_id");
            f.measurement_id = static_cast<uint16_t*>(o);
        }
        if (io.status) {
            e = stg.out(io.status, static_cast<size_t>(L.W) * 4, &o);
            if (e != cudaSuccess) return fail_cuda(e, "stage status");
            f.status = static_cast<uint32_t*>(o);
        }
        if (io.lut) {
            const void *fd = nullptr, *fo = nullptr;
            int fdt, fdev;
            size_t fh, fw;
            lut_view(io.lut, &fd, &fo, &fdt, &fh, &fw, &fdev);
            if (fh != L.H || fw != L.W) return fail(OB_INVALID_ARGUMENT, "unexpected image dimensions");
            if (fdev != device) return fail(OB_INVALID_ARGUMENT, "lut and stream are on different devices");
            if (lut && fdt != ldtype) return fail(OB_INVALID_ARGUMENT, "per-frame lut dtype differs from the call-level lut");
            if (!lut) ldtype = fdt;
            f.lut_dir = fd;
            f.lut_off = fo;
            f.lut_maps = maps_for(L, device, fd, fo, fdt);
            f.lut_an = lut_analytic(io.lut);
            if (!f.lut_maps && !f.lut_an) frame_maps_ok = false;
            if (!al16(fd) || !al16(fo)) vec_ok = false;
        }
        for (int r = 0; r < OB_MAX_RETURNS; ++r) {
            if (io.xyz[r]) {
                if (!lut && !io.lut) return fail(OB_INVALID_ARGUMENT, "xyz output requested without a lut");
                e = stg.out(io.xyz[r], n_px * 3 * (ldtype == OB_F64 ? 8 : 4), &o);
                if (e != cudaSuccess) return fail_cuda(e, "stage xyz");
                f.xyz[r] = o;
                any_xyz = true;
                if (!al16(o)) vec_ok = false;
            }
            if (io.range_destaggered[r]) {
                if (!shifts) return fail(OB_INVALID_ARGUMENT, "image height does not match shifts size");
                e = stg.out(io.range_destaggered[r], n_px * 4, &o);
                if (e != cudaSuccess) return fail_cuda(e, "stage range_destaggered");
                f.rd[r] = static_cast<uint32_t*>(o);
            }
        }
    }
    group_by_lut(hf);
    const void* fdev = nullptr;
    cudaError_t e = stream_table(s, 0, hf.data(), n_frames * sizeof(DecodeFrame), &fdev);
    if (e != cudaSuccess) return fail_cuda(e, "frame table upload");
    DecodeLaunch a;
    a.layout_host = &L;
    a.frames_dev = static_cast<const DecodeFrame*>(fdev);
    a.n_frames = static_cast<uint32_t>(n_frames);
    a.lut_dir = ldir;
    a.lut_off = loff;
    a.lut_dtype = ldtype;
    a.shift_host = shifts ? sh.data() : nullptr;
    a.vec_ok = vec_ok;
    a.lut_maps = maps_for(L, device, ldir, loff, ldtype);
    a.lut_an = lut ? lut_analytic(lut) : nullptr;
    a.frame_luts_have_maps = frame_maps_ok;
    a.any_xyz = any_xyz;
    e = launch_decode(a, device, st);
    if (e != cudaSuccess) return fail_cuda(e, "decode launch");
    e = stg.flush();
    if (e != cudaSuccess) return fail_cuda(e, "decode D2H");
    return OB_OK;
}

ob_status ob_decode_batch_run(const ob_decoder* dec, const ob_decode_batch* b, const ob_lut* lut,
                              const int32_t* shifts, size_t n_shifts, ob_stream* s) {
    if (!dec || !b || !s) return fail(OB_INVALID_ARGUMENT, "null pointer");
    if (b->n_frames == 0) return OB_OK;
    const DecodeLayout& L = dec->L;
    const int device = stream_device(s);
    if (device != dec->device) return fail(OB_INVALID_ARGUMENT, "decoder and stream are on different devices");
    ob_status rs = require_device(device);
    if (rs != OB_OK) return rs;
    if (!b->packets || b->n_slots == 0) return fail(OB_INVALID_ARGUMENT, "null packet buffer");
    if (b->packet_stride < L.packet_size)
        return fail(OB_INVALID_ARGUMENT, "packet_stride smaller than the lidar packet size");
    const void *ldir = nullptr, *loff = nullptr;
    int ldtype = OB_F32;
    if (lut) {
        size_t lh, lw;
        int ldev;
        lut_view(lut, &ldir, &loff, &ldtype, &lh, &lw, &ldev);
        if (lh != L.H || lw != L.W) return fail(OB_INVALID_ARGUMENT, "unexpected image dimensions");
        if (ldev != device) return fail(OB_INVALID_ARGUMENT, "lut and stream are on different devices");
    }
    std::vector<uint16_t> sh;
    if (shifts) {
        if (n_shifts != L.H) return fail(OB_INVALID_ARGUMENT, "image height does not match shifts size");
        if (L.H > static_cast<uint32_t>(kMaxRows))
            return fail(OB_INVALID_ARGUMENT, "fused destagger supports at most 512 rows");
        reduce_shifts(shifts, L.H, L.W, 0, sh);
    }
    const size_t F = b->n_frames;
    const size_t n_px = static_cast<size_t>(L.H) * L.W;
    cudaStream_t st = stream_handle(s);
    Staging stg(st);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    auto span = [&](size_t stride, size_t last) { return (F - 1) * stride + last; };
    bool vec_ok = !lut || (al16(ldir) && al16(loff));

    const void* dpk = nullptr;
    cudaError_t e = stg.in(b->packets,
                           span(b->packets_frame_stride, (b->n_slots - 1) * b->packet_stride + L.packet_size),
                           &dpk);
    if (e != cudaSuccess) return fail_cuda(e, "stage packets");
    void* dfields[OB_MAX_FIELDS] = {};
    for (uint32_t k = 0; k < L.n_fields; ++k) {
        if (!b->fields[k]) continue;
        e = stg.out(b->fields[k], span(b->field_frame_stride[k], n_px * L.fields[k].elem_size), &dfields[k]);
        if (e != cudaSuccess) return fail_cuda(e, "stage field output");
    }
    void *dts = nullptr, *dmid = nullptr, *dstat = nullptr;
    if (b->timestamp) e = stg.out(b->timestamp, span(b->timestamp_frame_stride, L.W * 8ull), &dts);
    if (e == cudaSuccess && b->measurement_id)
        e = stg.out(b->measurement_id, span(b->measurement_id_frame_stride, L.W * 2ull), &dmid);
    if (e == cudaSuccess && b->status) e = stg.out(b->status, span(b->status_frame_stride, L.W * 4ull), &dstat);
    if (e != cudaSuccess) return fail_cuda(e, "stage headers");
    std::vector<const void*> fl_dir, fl_off;
    if (b->frame_luts) {
        fl_dir.resize(F);
        fl_off.resize(F);
        for (size_t f = 0; f < F; ++f) {
            if (!b->frame_luts[f]) return fail(OB_INVALID_ARGUMENT, "null per-frame lut");
            int fdt, fdev;
            size_t fh, fw;
            lut_view(b->frame_luts[f], &fl_dir[f], &fl_off[f], &fdt, &fh, &fw, &fdev);
            if (fh != L.H || fw != L.W) return fail(OB_INVALID_ARGUMENT, "unexpected image dimensions");
            if (fdev != device) return fail(OB_INVALID_ARGUMENT, "lut and stream are on different devices");
            if ((lut || f > 0) && fdt != ldtype)
                return fail(OB_INVALID_ARGUMENT, "per-frame lut dtype differs");
            ldtype = fdt;
            if (!al16(fl_dir[f]) || !al16(fl_off[f])) vec_ok = false;
        }
    }
    const size_t esz2 = ldtype == OB_F64 ? 8 : 4;
    void* dxyz[OB_MAX_RETURNS] = {};
    void* drd[OB_MAX_RETURNS] = {};
    for (int r = 0; r < OB_MAX_RETURNS; ++r) {
        if (b->xyz[r]) {
            if (!lut && !b->frame_luts) return fail(OB_INVALID_ARGUMENT, "xyz output requested without a lut");
            e = stg.out(b->xyz[r], span(b->xyz_frame_stride, n_px * 3 * esz2), &dxyz[r]);
            if (e != cudaSuccess) return fail_cuda(e, "stage xyz");
            if (!al16(dxyz[r]) || b->xyz_frame_stride % 16) vec_ok = false;
        }
        if (b->range_destaggered[r]) {
            if (!shifts) return fail(OB_INVALID_ARGUMENT, "image height does not match shifts size");
            e = stg.out(b->range_destaggered[r], span(b->rd_frame_stride, n_px * 4), &drd[r]);
            if (e != cudaSuccess) return fail_cuda(e, "stage range_destaggered");
        }
    }
    std::vector<DecodeFrame> hf(F);
    bool frame_maps_ok = true;
    const bool bulk_ok = al16(dpk) && b->packet_stride % 16 == 0 && L.packet_size % 16 == 0 &&
                         b->packets_frame_stride % 16 == 0;
    for (size_t f = 0; f < F; ++f) {
        DecodeFrame& d = hf[f];
        std::memset(&d, 0, sizeof(d));
        d.packets = static_cast<const uint8_t*>(dpk) + f * b->packets_frame_stride;
        d.packet_stride = b->packet_stride;
        d.n_slots = static_cast<uint32_t>(b->n_slots);
        d.flags = 1u | (bulk_ok ? 2u : 0u);
        bool allf = L.n_fields > 0;
        for (uint32_t k = 0; k < L.n_fields; ++k) {
            if (dfields[k]) d.fields[k] = static_cast<uint8_t*>(dfields[k]) + f * b->field_frame_stride[k];
            allf = allf && dfields[k] != nullptr;
        }
        if (allf) d.flags |= 4u;  // every decoder field has an output image
        if (dts) d.timestamp = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(dts) + f * b->timestamp_frame_stride);
        if (dmid) d.measurement_id = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(dmid) + f * b->measurement_id_frame_stride);
        if (dstat) d.status = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(dstat) + f * b->status_frame_stride);
        if (!fl_dir.empty()) {
            d.lut_dir = fl_dir[f];
            d.lut_off = fl_off[f];
            d.lut_maps = maps_for(L, device, fl_dir[f], fl_off[f], ldtype);
            d.lut_an = lut_analytic(b->frame_luts[f]);
            if (!d.lut_maps && !d.lut_an) frame_maps_ok = false;
        }
        for (int r = 0; r < OB_MAX_RETURNS; ++r) {
            if (dxyz[r]) d.xyz[r] = static_cast<uint8_t*>(dxyz[r]) + f * b->xyz_frame_stride;
            if (drd[r]) d.rd[r] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(drd[r]) + f * b->rd_frame_stride);
        }
    }
    group_by_lut(hf);
    const void* fdev = nullptr;
    e = stream_table(s, 0, hf.data(), F * sizeof(DecodeFrame), &fdev);
    if (e != cudaSuccess) return fail_cuda(e, "frame table upload");
    DecodeLaunch a;
    a.layout_host = &L;
    a.frames_dev = static_cast<const DecodeFrame*>(fdev);
    a.n_frames = static_cast<uint32_t>(F);
    a.lut_dir = ldir;
    a.lut_off = loff;
    a.lut_dtype = ldtype;
    a.shift_host = shifts ? sh.data() : nullptr;
    a.vec_ok = vec_ok;
    a.lut_maps = maps_for(L, device, ldir, loff, ldtype);
    a.lut_an = lut ? lut_analytic(lut) : nullptr;
    a.frame_luts_have_maps = frame_maps_ok;
    a.any_xyz = dxyz[0] != nullptr || dxyz[1] != nullptr;
    a.xyz_base[0] = dxyz[0];
    a.xyz_base[1] = dxyz[1];
    a.xyz_frame_stride = b->xyz_frame_stride;
    e = launch_decode(a, device, st);
    if (e != cudaSuccess) return fail_cuda(e, "decode launch");
    e = stg.flush();
    if (e != cudaSuccess) return fail_cuda(e, "decode D2H");
    return OB_OK;
}


// ---------------------------------------------------------------------------------------------
// decode job
// ---------------------------------------------------------------------------------------------
struct ob_decode_job {
    const ob_decoder* dec;
    ob_stream* s;
    cudaStream_t st;
    int device;
    size_t stride;             // bytes per packet slot on the device (multiple of 16)
    uint8_t* d_pk{nullptr};    // packet slots
    size_t cap_slots{0};
    size_t up_slots{0};        // highest uploaded slot + 1
    uint8_t* d_out{nullptr};   // output slab (fields, headers, xyz, destaggered ranges)
    size_t out_bytes{0};
    int32_t* d_colsrc{nullptr};
    int32_t* h_colsrc{nullptr};     // pinned
    DecodeFrame* d_frame{nullptr};
    DecodeFrame* h_frame{nullptr};  // pinned
    cudaEvent_t ev_up{nullptr}, ev_done{nullptr};
    bool up_pending{false}, busy{false};
};

static cudaError_t job_reserve(ob_decode_job* j, size_t slots) {
    if (slots <= j->cap_slots) return cudaSuccess;
    const size_t cap = std::max(slots, j->cap_slots ? j->cap_slots * 2 : static_cast<size_t>(16));
    uint8_t* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap * j->stride + 16);
    if (e != cudaSuccess) return e;
    if (j->d_pk) {  // rare: more packets than the frame was sized for (duplicates, retransmits)
        e = cudaStreamSynchronize(j->st);
        if (e == cudaSuccess && j->up_slots)
            e = cudaMemcpy(p, j->d_pk, j->up_slots * j->stride, cudaMemcpyDeviceToDevice);
        cudaFree(j->d_pk);
        if (e != cudaSuccess) {
            cudaFree(p);
            j->d_pk = nullptr;
            j->cap_slots = 0;
            return e;
        }
    }
    j->d_pk = p;
    j->cap_slots = cap;
    return cudaSuccess;
}

ob_status ob_decode_job_create(const ob_decoder* dec, size_t reserve_slots, ob_stream* s,
                               ob_decode_job** out) {
    if (!dec || !s || !out) return fail(OB_INVALID_ARGUMENT, "null pointer");
    const int device = stream_device(s);
    if (device != dec->device) return fail(OB_INVALID_ARGUMENT, "decoder and stream are on different devices");
    ob_status rs = require_device(device);
    if (rs != OB_OK) return rs;
    std::unique_ptr<ob_decode_job> j(new ob_decode_job);
    j->dec = dec;
    j->s = s;
    j->st = stream_handle(s);
    j->device = device;
    j->stride = (static_cast<size_t>(dec->L.packet_size) + 15) & ~static_cast<size_t>(15);
    cudaError_t e = cudaEventCreateWithFlags(&j->ev_up, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&j->d_colsrc, static_cast<size_t>(dec->L.W) * 4);
    if (e == cudaSuccess) e = cudaHostAlloc(&j->h_colsrc, static_cast<size_t>(dec->L.W) * 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&j->d_frame, sizeof(DecodeFrame));
    if (e == cudaSuccess) e = cudaHostAlloc(&j->h_frame, sizeof(DecodeFrame), cudaHostAllocDefault);
    if (e == cudaSuccess && reserve_slots) e = job_reserve(j.get(), reserve_slots);
    if (e != cudaSuccess) {
        ob_decode_job_destroy(j.release());
        return fail_cuda(e, "decode job allocation");
    }
    *out = j.release();
    return OB_OK;
}

ob_status ob_decode_job_destroy(ob_decode_job* j) {
    if (!j) return OB_OK;
    cudaSetDevice(j->device);
    cudaStreamSynchronize(j->st);
    if (j->ev_up) cudaEventDestroy(j->ev_up);
    if (j->ev_done) cudaEventDestroy(j->ev_done);
    cudaFree(j->d_pk);
    cudaFree(j->d_out);
    cudaFree(j->d_colsrc);
    cudaFree(j->d_frame);
    if (j->h_colsrc) cudaFreeHost(j->h_colsrc);
    if (j->h_frame) cudaFreeHost(j->h_frame);
    delete j;
    return OB_OK;
}

int ob_decode_job_busy(const ob_decode_job* j) { return j && j->busy ? 1 : 0; }

ob_status ob_decode_job_upload(ob_decode_job* j, const uint8_t* src, size_t src_stride,
                               size_t first_slot, size_t count) {
    if (!j || (count && !src)) return fail(OB_INVALID_ARGUMENT, "null pointer");
    if (count == 0) return OB_OK;
    const size_t psize = j->dec->L.packet_size;
    if (count > 1 && src_stride < psize)
        return fail(OB_INVALID_ARGUMENT, "packet_stride smaller than the lidar packet size");
    if (first_slot + count > (1u << 20)) return fail(OB_INVALID_ARGUMENT, "too many packet slots");
    ob_status rs = require_device(j->device);
    if (rs != OB_OK) return rs;
    if (j->busy) {  // the previous frame still reads the slots
        rs = ob_decode_job_wait(j);
        if (rs != OB_OK) return rs;
    }
    cudaError_t e = job_reserve(j, first_slot + count);
    if (e != cudaSuccess) return fail_cuda(e, "decode job packet slots");
    uint8_t* dst = j->d_pk + first_slot * j->stride;
    if (count == 1 || src_stride == j->stride)
        e = cudaMemcpyAsync(dst, src, (count - 1) * j->stride + psize, cudaMemcpyDefault, j->st);
    else
        e = cudaMemcpy2DAsync(dst, j->stride, src, src_stride, psize, count, cudaMemcpyDefault, j->st);
    if (e != cudaSuccess) return fail_cuda(e, "packet upload");
    e = cudaEventRecord(j->ev_up, j->st);
    if (e != cudaSuccess) return fail_cuda(e, "packet upload event");
    j->up_pending = true;
    // an upload at slot 0 begins a new frame: slots of the previous one are no longer valid
    j->up_slots = first_slot == 0 ? count : std::max(j->up_slots, first_slot + count);
    return OB_OK;
}

ob_status ob_decode_job_uploads_done(ob_decode_job* j) {
    if (!j) return fail(OB_INVALID_ARGUMENT, "null pointer");
    if (!j->up_pending) return OB_OK;
    cudaError_t e = cudaEventSynchronize(j->ev_up);
    j->up_pending = false;
    if (e != cudaSuccess) return fail_cuda(e, "packet upload");
    return OB_OK;
}

ob_status ob_decode_job_wait(ob_decode_job* j) {
    if (!j) return fail(OB_INVALID_ARGUMENT, "null pointer");
    if (!j->busy) return OB_OK;
    cudaError_t e = cudaEventSynchronize(j->ev_done);
    j->busy = false;
    j->up_pending = false;
    if (e != cudaSuccess) return fail_cuda(e, "decode job");
    return OB_OK;
}

ob_status ob_decode_job_submit(ob_decode_job* j, const ob_decode_io* io, const ob_lut* lut,
                               const int32_t* shifts, size_t n_shifts) {
    if (!j || !io) return fail(OB_INVALID_ARGUMENT, "null pointer");
    const DecodeLayout& L = j->dec->L;
    ob_status rs = require_device(j->device);
    if (rs != OB_OK) return rs;
    if (io->n_slots > j->up_slots) return fail(OB_INVALID_ARGUMENT, "n_slots exceeds the uploaded packet slots");
    const ob_lut* use_lut = io->lut ? io->lut : lut;
    const void *ldir = nullptr, *loff = nullptr;
    int ldtype = OB_F32;
    if (use_lut) {
        size_t lh, lw;
        int ldev;
        lut_view(use_lut, &ldir, &loff, &ldtype, &lh, &lw, &ldev);
        if (lh != L.H || lw != L.W) return fail(OB_INVALID_ARGUMENT, "unexpected image dimensions");
        if (ldev != j->device) return fail(OB_INVALID_ARGUMENT, "lut and stream are on different devices");
    }
    std::vector<uint16_t> sh;
    if (shifts) {
        if (n_shifts != L.H) return fail(OB_INVALID_ARGUMENT, "image height does not match shifts size");
        if (L.H > static_cast<uint32_t>(kMaxRows))
            return fail(OB_INVALID_ARGUMENT, "fused destagger supports at most 512 rows");
        reduce_shifts(shifts, L.H, L.W, 0, sh);
    }
    for (int r = 0; r < OB_MAX_RETURNS; ++r) {
        if (io->xyz[r] && !use_lut) return fail(OB_INVALID_ARGUMENT, "xyz output requested without a lut");
        if (io->range_destaggered[r] && !shifts)
            return fail(OB_INVALID_ARGUMENT, "image height does not match shifts size");
    }
    if (j->busy) {  // the previous submission still owns h_frame / h_colsrc / the slab
        rs = ob_decode_job_wait(j);
        if (rs != OB_OK) return rs;
    }

    // ---- outputs: device pointers in place, host pointers through the slab + D2H ----
    const size_t n_px = static_cast<size_t>(L.H) * L.W;
    struct Out {
        void* user;
        size_t bytes;
        void** slot;  // where the device pointer goes
        bool host;
        size_t off;
    };
    DecodeFrame f;
    std::memset(&f, 0, sizeof(f));
    Out outs[OB_MAX_FIELDS + 3 + 2 * OB_MAX_RETURNS];
    size_t n_out = 0;
    void* fld[OB_MAX_FIELDS] = {};
    void *ts = nullptr, *mid = nullptr, *stt = nullptr, *xyz[OB_MAX_RETURNS] = {}, *rd[OB_MAX_RETURNS] = {};
    auto add = [&](void* user, size_t bytes, void** slot) {
        if (user) outs[n_out++] = Out{user, bytes, slot, false, 0};
    };
    for (uint32_t k = 0; k < L.n_fields; ++k) add(io->fields[k], n_px * L.fields[k].elem_size, &fld[k]);
    add(io->timestamp, static_cast<size_t>(L.W) * 8, &ts);
    add(io->measurement_id, static_cast<size_t>(L.W) * 2, &mid);
    add(io->status, static_cast<size_t>(L.W) * 4, &stt);
    for (int r = 0; r < OB_MAX_RETURNS; ++r) {
        add(io->xyz[r], n_px * 3 * (ldtype == OB_F64 ? 8 : 4), &xyz[r]);
        add(io->range_destaggered[r], n_px * 4, &rd[r]);
    }
    // host outputs take slab space in ADDRESS order; buffers that are adjacent in host memory
    // (HostBuffer::carve) stay adjacent in the slab, so their D2H is one copy
    size_t order[OB_MAX_FIELDS + 3 + 2 * OB_MAX_RETURNS];
    size_t n_host = 0;
    for (size_t i = 0; i < n_out; ++i) {
        outs[i].host = !is_device_ptr(outs[i].user);
        if (outs[i].host) order[n_host++] = i;
    }
    std::sort(order, order + n_host, [&](size_t a, size_t b) {
        return reinterpret_cast<uintptr_t>(outs[a].user) < reinterpret_cast<uintptr_t>(outs[b].user);
    });
    size_t need = 0;
    for (size_t k = 0; k < n_host; ++k) {
        Out& o = outs[order[k]];
        const bool adjacent = k > 0 && static_cast<uint8_t*>(outs[order[k - 1]].user) + outs[order[k - 1]].bytes ==
                                           static_cast<uint8_t*>(o.user) &&
                              (need % 16) == 0;
        if (!adjacent) need = (need + 255) & ~static_cast<size_t>(255);
        o.off = need;
        need += o.bytes;
    }
    cudaError_t e = cudaSuccess;
    if (need > j->out_bytes) {  // job is idle here
        cudaFree(j->d_out);
        j->d_out = nullptr;
        j->out_bytes = 0;
        e = cudaMalloc(&j->d_out, need);
        if (e != cudaSuccess) return fail_cuda(e, "decode job output slab");
        j->out_bytes = need;
    }
    for (size_t i = 0; i < n_out; ++i) *outs[i].slot = outs[i].host ? j->d_out + outs[i].off : outs[i].user;

    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    bool vec_ok = !use_lut || (al16(ldir) && al16(loff));
    f.packets = j->d_pk;
    f.packet_stride = j->stride;
    f.n_slots = static_cast<uint32_t>(io->n_slots);
    f.flags = 0;
    if (!io->col_src) f.flags |= 1u;
    if (L.packet_size % 16 == 0) f.flags |= 2u;  // slots are 16-byte aligned by construction
    if (io->col_src) {
        std::memcpy(j->h_colsrc, io->col_src, static_cast<size_t>(L.W) * 4);
        e = cudaMemcpyAsync(j->d_colsrc, j->h_colsrc, static_cast<size_t>(L.W) * 4, cudaMemcpyHostToDevice, j->st);
        if (e != cudaSuccess) return fail_cuda(e, "stage column map");
        f.col_src = j->d_colsrc;
    }
    {
        bool allf = L.n_fields > 0;
        for (uint32_t k = 0; k < L.n_fields; ++k) {
            f.fields[k] = fld[k];
            allf = allf && fld[k] != nullptr;
        }
        if (allf) f.flags |= 4u;  // every decoder field has an output image
    }
    f.timestamp = static_cast<uint64_t*>(ts);
    f.measurement_id = static_cast<uint16_t*>(mid);
    f.status = static_cast<uint32_t*>(stt);
    for (int r = 0; r < OB_MAX_RETURNS; ++r) {
        f.xyz[r] = xyz[r];
        f.rd[r] = static_cast<uint32_t*>(rd[r]);
        if (xyz[r] && !al16(xyz[r])) vec_ok = false;
    }
    *j->h_frame = f;
    e = cudaMemcpyAsync(j->d_frame, j->h_frame, sizeof(DecodeFrame), cudaMemcpyHostToDevice, j->st);
    if (e != cudaSuccess) return fail_cuda(e, "frame table upload");
    DecodeLaunch a;
    a.layout_host = &L;
    a.frames_dev = j->d_frame;
    a.n_frames = 1;
    a.lut_dir = ldir;
    a.lut_off = loff;
    a.lut_dtype = ldtype;
    a.shift_host = shifts ? sh.data() : nullptr;
    a.vec_ok = vec_ok;
    a.lut_maps = maps_for(L, j->device, ldir, loff, ldtype);
    a.lut_an = use_lut ? lut_analytic(use_lut) : nullptr;
    a.any_xyz = xyz[0] != nullptr || xyz[1] != nullptr;
    a.xyz_base[0] = xyz[0];  // a one-frame batch
    a.xyz_base[1] = xyz[1];
    a.xyz_frame_stride = static_cast<unsigned long long>(n_px) * 3 * (ldtype == OB_F64 ? 8 : 4);
    e = launch_decode(a, j->device, j->st);
    if (e != cudaSuccess) return fail_cuda(e, "decode launch");
    for (size_t k = 0; k < n_host;) {  // one D2H per run of outputs contiguous on both sides
        const Out& first = outs[order[k]];
        size_t bytes = first.bytes, m = k + 1;
        while (m < n_host && outs[order[m]].off == first.off + bytes &&
               static_cast<uint8_t*>(outs[order[m]].user) == static_cast<uint8_t*>(first.user) + bytes) {
            bytes += outs[order[m]].bytes;
            ++m;
        }
        e = cudaMemcpyAsync(first.user, j->d_out + first.off, bytes, cudaMemcpyDeviceToHost, j->st);
        if (e != cudaSuccess) return fail_cuda(e, "decode D2H");
        k = m;
    }
    e = cudaEventRecord(j->ev_done, j->st);
    if (e != cudaSuccess) return fail_cuda(e, "decode job event");
    j->busy = true;
    return OB_OK;
}

ob_status ob_encode_frames(const ob_decoder* dec, const ob_encode_io* frames, size_t n_frames, int with_crc,
                           ob_stream* s) {
    if (!dec || !s || (n_frames && !frames)) return fail(OB_INVALID_ARGUMENT, "null pointer");
    if (n_frames == 0) return OB_OK;
    const DecodeLayout& L = dec->L;
    const int device = stream_device(s);
    if (device != dec->device) return fail(OB_INVALID_ARGUMENT, "decoder and stream are on different devices");
    if (L.W % L.cpp != 0)
        return fail(OB_INVALID_ARGUMENT, "Mismatch between expected number of packets and PacketFormat.columns_per_packet");
    if (with_crc && (L.packet_size % 4 != 0 || L.packet_size < 8))
        return fail(OB_INVALID_ARGUMENT, "packet size must be a multiple of 4 for the CRC64 footer");
    ob_status rs = require_device(device);
    if (rs != OB_OK) return rs;
    cudaStream_t st = stream_handle(s);
    Staging stg(st);
    const size_t n_px = static_cast<size_t>(L.H) * L.W, n_pk = L.W / L.cpp;
    std::vector<EncodeFrame> hf(n_frames);
    for (size_t i = 0; i < n_frames; ++i) {
        const ob_encode_io& io = frames[i];
        EncodeFrame& f = hf[i];
        std::memset(&f, 0, sizeof(f));
        if (!io.packets) return fail(OB_INVALID_ARGUMENT, "null packet buffer");
        if (io.packet_stride < L.packet_size)
            return fail(OB_INVALID_ARGUMENT, "packet_stride smaller than the lidar packet size");
        if (io.packet_headers && io.packet_header_bytes < L.packet_header_size)
            return fail(OB_INVALID_ARGUMENT, "packet_header_bytes smaller than the packet header");
        const void* d = nullptr;
        cudaError_t e = cudaSuccess;
        for (uint32_t k = 0; k < L.n_fields && e == cudaSuccess; ++k) {
            if (!io.fields[k]) continue;
            e = stg.in(io.fields[k], n_px * L.fields[k].elem_size, &d);
            f.fields[k] = d;
        }
        if (e == cudaSuccess && io.timestamp) {
            e = stg.in(io.timestamp, static_cast<size_t>(L.W) * 8, &d);
            f.timestamp = static_cast<const uint64_t*>(d);
        }
        if (e == cudaSuccess && io.status) {
            e = stg.in(io.status, static_cast<size_t>(L.W) * 4, &d);
            f.status = static_cast<const uint32_t*>(d);
        }
        if (e == cudaSuccess && io.packet_headers) {
            e = stg.in(io.packet_headers, n_pk * io.packet_header_bytes, &d);
            f.packet_headers = static_cast<const uint8_t*>(d);
            f.header_bytes = static_cast<uint32_t>(io.packet_header_bytes);
        }
        void* o = nullptr;
        if (e == cudaSuccess) e = stg.out(io.packets, (n_pk - 1) * io.packet_stride + L.packet_size, &o);
        if (e != cudaSuccess) return fail_cuda(e, "stage encode buffers");
        f.packets = static_cast<uint8_t*>(o);
        f.packet_stride = io.packet_stride;
    }
    const void* fdev = nullptr;
    cudaError_t e = stream_table(s, 1, hf.data(), n_frames * sizeof(EncodeFrame), &fdev);
    if (e != cudaSuccess) return fail_cuda(e, "frame table upload");
    e = launch_encode(L, static_cast<const EncodeFrame*>(fdev), static_cast<uint32_t>(n_frames), with_crc != 0, device, st);
    if (e != cudaSuccess) return fail_cuda(e, "encode launch");
    e = stg.flush();
    if (e != cudaSuccess) return fail_cuda(e, "encode D2H");
    return OB_OK;
}

int ob_pointer_kind(const void* p) {
    if (!p) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (at.type == cudaMemoryTypeHost) return 1;
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return 2;
    return 0;
}

int ob_pointer_host_readable(const void* p) {
    if (!p) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return at.type == cudaMemoryTypeDevice ? 0 : 1;
}

}  // extern "C"
