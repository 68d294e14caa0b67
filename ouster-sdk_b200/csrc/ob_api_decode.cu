// ob_api_decode.cu -- C-ABI glue of the packet-decode path (ob_decoder_*, ob_decode_frames).
#include <cstring>
#include <vector>

#include "ob_api_common.h"

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
    bool vec_ok = true;
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
            if (!al16(fd) || !al16(fo)) vec_ok = false;
        }
        for (int r = 0; r < OB_MAX_RETURNS; ++r) {
            if (io.xyz[r]) {
                if (!lut && !io.lut) return fail(OB_INVALID_ARGUMENT, "xyz output requested without a lut");
                e = stg.out(io.xyz[r], n_px * 3 * (ldtype == OB_F64 ? 8 : 4), &o);
                if (e != cudaSuccess) return fail_cuda(e, "stage xyz");
                f.xyz[r] = o;
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
    void* fdev = nullptr;
    cudaError_t e = stg.scratch(n_frames * sizeof(DecodeFrame), &fdev);
    if (e != cudaSuccess) return fail_cuda(e, "frame table alloc");
    e = cudaMemcpyAsync(fdev, hf.data(), n_frames * sizeof(DecodeFrame), cudaMemcpyHostToDevice, st);
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
    const bool bulk_ok = al16(dpk) && b->packet_stride % 16 == 0 && L.packet_size % 16 == 0 &&
                         b->packets_frame_stride % 16 == 0;
    for (size_t f = 0; f < F; ++f) {
        DecodeFrame& d = hf[f];
        std::memset(&d, 0, sizeof(d));
        d.packets = static_cast<const uint8_t*>(dpk) + f * b->packets_frame_stride;
        d.packet_stride = b->packet_stride;
        d.n_slots = static_cast<uint32_t>(b->n_slots);
        d.flags = 1u | (bulk_ok ? 2u : 0u);
        for (uint32_t k = 0; k < L.n_fields; ++k)
            if (dfields[k]) d.fields[k] = static_cast<uint8_t*>(dfields[k]) + f * b->field_frame_stride[k];
        if (dts) d.timestamp = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(dts) + f * b->timestamp_frame_stride);
        if (dmid) d.measurement_id = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(dmid) + f * b->measurement_id_frame_stride);
        if (dstat) d.status = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(dstat) + f * b->status_frame_stride);
        if (!fl_dir.empty()) {
            d.lut_dir = fl_dir[f];
            d.lut_off = fl_off[f];
        }
        for (int r = 0; r < OB_MAX_RETURNS; ++r) {
            if (dxyz[r]) d.xyz[r] = static_cast<uint8_t*>(dxyz[r]) + f * b->xyz_frame_stride;
            if (drd[r]) d.rd[r] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(drd[r]) + f * b->rd_frame_stride);
        }
    }
    void* fdev = nullptr;
    e = stg.scratch(F * sizeof(DecodeFrame), &fdev);
    if (e != cudaSuccess) return fail_cuda(e, "frame table alloc");
    e = cudaMemcpyAsync(fdev, hf.data(), F * sizeof(DecodeFrame), cudaMemcpyHostToDevice, st);
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
    e = launch_decode(a, device, st);
    if (e != cudaSuccess) return fail_cuda(e, "decode launch");
    e = stg.flush();
    if (e != cudaSuccess) return fail_cuda(e, "decode D2H");
    return OB_OK;
}

}  // extern "C"
