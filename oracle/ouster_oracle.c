/*
 * ouster_oracle.c -- CPU oracle (plain C restatement) of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see ouster_oracle.h.  Each function cites the
 * reference file:line it follows (paths relative to /root/reference).
 * Build: see oracle/Makefile (-O3 -DNDEBUG -ffp-contract=off, no -march; mirrors
 * cmake/DefaultBuildType.cmake:2-5 of the reference, which has no FMA).
 */
#include "ouster_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* FieldDecodeInfo                                                            */
/* ------------------------------------------------------------------------- */

/* field_info() factory -- ouster_core/src/parsing.cpp:57-122 */
int orc_field_info_make(size_t bit_start, size_t bit_size, size_t upshift, size_t max_length,
                        size_t num_elements, orc_field_info* out) {
    orc_field_info fi;
    memset(&fi, 0, sizeof(fi));
    size_t needs_bits = bit_size + upshift;
    if (needs_bits > 64) return -1;

    fi.offset = bit_start / 8;
    bit_start %= 8;
    for (size_t i = bit_start; i < bit_start + bit_size; ++i) fi.mask |= (uint64_t)1 << i;
    fi.shift = (int)bit_start - (int)upshift;
    fi.num_elements = (int)num_elements;

    size_t size_bytes = needs_bits / 8 + ((needs_bits % 8) ? 1 : 0);
    size_bytes /= num_elements;
    switch (size_bytes) {
        case 1: fi.ty_tag = ORC_UINT8; break;
        case 2: fi.ty_tag = ORC_UINT16; break;
        case 3: case 4: fi.ty_tag = ORC_UINT32; break;
        case 5: case 6: case 7: case 8: fi.ty_tag = ORC_UINT64; break;
        default: fi.ty_tag = ORC_VOID;
    }
    if (max_length > 0) {
        if (fi.offset + size_bytes > max_length) return -1;
        int needed = (int)fi.offset + 8 - (int)max_length;
        if (needed > 0) {
            fi.offset -= (size_t)needed;
            fi.mask <<= needed * 8;
            fi.shift += needed * 8;
        }
    }
    *out = fi;
    return 0;
}

/* FieldDecodeInfo::get<T> -- field_decode_info.h:41-54 (caller truncates to sizeof(T)) */
uint64_t orc_field_get(const orc_field_info* fi, const uint8_t* buffer) {
    uint64_t word;
    memcpy(&word, buffer + fi->offset, 8); /* unaligned little-endian 8-byte load */
    word &= fi->mask;
    if (fi->shift > 0) {
        word >>= fi->shift;
    } else if (fi->shift < 0) {
        word <<= -fi->shift;
    }
    return word;
}

/* FieldDecodeInfo::set<T> -- field_decode_info.h:64-78 (value already widened to 64 bit) */
void orc_field_set(const orc_field_info* fi, uint8_t* buffer, uint64_t value) {
    uint64_t word = value;
    if (fi->shift > 0) word <<= fi->shift;
    if (fi->shift < 0) word >>= -fi->shift;
    word &= fi->mask;
    uint64_t cur;
    memcpy(&cur, buffer + fi->offset, 8);
    cur &= ~fi->mask;
    cur |= word;
    memcpy(buffer + fi->offset, &cur, 8);
}

size_t orc_type_size(int ty) {
    switch (ty) {
        case ORC_UINT8: case ORC_INT8: case ORC_CHAR: return 1;
        case ORC_UINT16: case ORC_INT16: case ORC_FLOAT16: return 2;
        case ORC_UINT32: case ORC_INT32: case ORC_FLOAT32: return 4;
        case ORC_UINT64: case ORC_INT64: case ORC_FLOAT64: return 8;
        default: return 0;
    }
}

static uint64_t type_mask(int ty) {
    switch (orc_type_size(ty)) {
        case 1: return 0xffull;
        case 2: return 0xffffull;
        case 4: return 0xffffffffull;
        case 8: return ~0ull;
        default: return 0;
    }
}

/* impl::get_value_mask -- parsing.cpp:139-156 */
uint64_t orc_value_mask(const orc_field_info* fi) {
    uint64_t tm = type_mask(fi->ty_tag);
    uint64_t m = fi->mask;
    if (m == 0) m = tm;
    if (fi->shift > 0) m >>= fi->shift;
    if (fi->shift < 0) m <<= -fi->shift;
    return m & tm;
}

/* ------------------------------------------------------------------------- */
/* Profile tables -- parsing.cpp:170-363                                      */
/* ------------------------------------------------------------------------- */

typedef struct { const char* name; unsigned bit, size, up, nel; } spec_t;

static const spec_t T_LEGACY[] = {
    {"RANGE", 0, 20, 0, 1}, {"FLAGS", 28, 4, 0, 1}, {"REFLECTIVITY", 32, 8, 0, 1},
    {"SIGNAL", 48, 16, 0, 1}, {"NEAR_IR", 64, 16, 0, 1}, {"RAW32_WORD1", 0, 32, 0, 1},
    {"RAW32_WORD2", 32, 32, 0, 1}, {"RAW32_WORD3", 64, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_LB[] = {
    {"RANGE", 0, 15, 3, 1}, {"FLAGS", 15, 1, 0, 1}, {"REFLECTIVITY", 16, 8, 0, 1},
    {"NEAR_IR", 24, 8, 4, 1}, {"RAW32_WORD1", 0, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_LB_WIN[] = {
    {"RANGE", 0, 15, 3, 1}, {"FLAGS", 15, 1, 0, 1}, {"REFLECTIVITY", 16, 8, 0, 1},
    {"WINDOW", 24, 8, 0, 1}, {"RAW32_WORD1", 0, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_RGB[] = {
    {"RANGE", 0, 19, 0, 1}, {"FLAGS", 19, 5, 0, 1}, {"REFLECTIVITY", 24, 8, 0, 1},
    {"SIGNAL", 32, 16, 0, 1}, {"NEAR_IR", 48, 16, 0, 1}, {"R", 64, 16, 0, 1},
    {"G", 80, 16, 0, 1}, {"B", 96, 16, 0, 1}, {"RGB", 64, 48, 0, 3},
    {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1}, {"RAW32_WORD3", 64, 32, 0, 1},
    {"RAW32_WORD4", 96, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_DUAL_RGB[] = {
    {"RANGE", 0, 19, 0, 1}, {"FLAGS", 19, 5, 0, 1}, {"REFLECTIVITY", 24, 8, 0, 1},
    {"RANGE2", 32, 19, 0, 1}, {"FLAGS2", 51, 5, 0, 1}, {"REFLECTIVITY2", 56, 8, 0, 1},
    {"SIGNAL", 64, 16, 0, 1}, {"SIGNAL2", 80, 16, 0, 1}, {"NEAR_IR", 96, 16, 0, 1},
    {"R", 112, 16, 0, 1}, {"G", 128, 16, 0, 1}, {"B", 144, 16, 0, 1}, {"RGB", 112, 48, 0, 3},
    {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1}, {"RAW32_WORD3", 64, 32, 0, 1},
    {"RAW32_WORD4", 96, 32, 0, 1}, {"RAW32_WORD5", 128, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_DUAL[] = {
    {"RANGE", 0, 19, 0, 1}, {"FLAGS", 19, 5, 0, 1}, {"REFLECTIVITY", 24, 8, 0, 1},
    {"RANGE2", 32, 19, 0, 1}, {"FLAGS2", 51, 5, 0, 1}, {"REFLECTIVITY2", 56, 8, 0, 1},
    {"SIGNAL", 64, 16, 0, 1}, {"SIGNAL2", 80, 16, 0, 1}, {"NEAR_IR", 96, 16, 0, 1},
    {"WINDOW", 120, 8, 0, 1}, {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1},
    {"RAW32_WORD3", 64, 32, 0, 1}, {"RAW32_WORD4", 96, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_SINGLE[] = {
    {"RANGE", 0, 19, 0, 1}, {"FLAGS", 19, 5, 0, 1}, {"REFLECTIVITY", 32, 8, 0, 1},
    {"SIGNAL", 48, 16, 0, 1}, {"NEAR_IR", 64, 16, 0, 1}, {"WINDOW", 88, 8, 0, 1},
    {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1}, {"RAW32_WORD3", 64, 32, 0, 1},
    {0, 0, 0, 0, 0}};
static const spec_t T_FIVE[] = {
    {"RANGE", 0, 19, 0, 1}, {"FLAGS", 19, 5, 0, 1}, {"REFLECTIVITY", 24, 8, 0, 1},
    {"RANGE2", 32, 19, 0, 1}, {"FLAGS2", 51, 5, 0, 1}, {"REFLECTIVITY2", 56, 8, 0, 1},
    {"SIGNAL", 64, 16, 0, 1}, {"SIGNAL2", 80, 16, 0, 1}, {"NEAR_IR", 96, 16, 0, 1},
    {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1}, {"RAW32_WORD3", 64, 32, 0, 1},
    {"RAW32_WORD4", 96, 32, 0, 1}, {"RAW32_WORD5", 128, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_ZM_LB[] = {
    {"RANGE", 0, 15, 3, 1}, {"FLAGS", 15, 1, 0, 1}, {"REFLECTIVITY", 16, 8, 0, 1},
    {"NEAR_IR", 24, 8, 4, 1}, {"ZONE_MASK", 32, 16, 0, 1}, {"WINDOW", 48, 8, 0, 1},
    {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_ZM_SINGLE[] = {
    {"RANGE", 0, 19, 0, 1}, {"FLAGS", 19, 5, 0, 1}, {"REFLECTIVITY", 32, 8, 0, 1},
    {"WINDOW", 40, 8, 0, 1}, {"SIGNAL", 48, 16, 0, 1}, {"NEAR_IR", 64, 16, 0, 1},
    {"ZONE_MASK", 80, 16, 0, 1}, {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1},
    {"RAW32_WORD3", 64, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_DUAL_LB[] = {
    {"RANGE", 0, 15, 3, 1}, {"FLAGS", 15, 1, 0, 1}, {"REFLECTIVITY", 16, 8, 0, 1},
    {"NEAR_IR", 24, 8, 4, 1}, {"RANGE2", 32, 15, 3, 1}, {"FLAGS2", 47, 1, 0, 1},
    {"REFLECTIVITY2", 48, 8, 0, 1}, {"WINDOW", 56, 8, 0, 1}, {"RAW32_WORD1", 0, 32, 0, 1},
    {"RAW32_WORD2", 32, 32, 0, 1}, {0, 0, 0, 0, 0}};
static const spec_t T_DUAL_ZONE[] = {
    {"RANGE", 0, 19, 0, 1}, {"FLAGS", 19, 5, 0, 1}, {"REFLECTIVITY", 24, 8, 0, 1},
    {"RANGE2", 32, 19, 0, 1}, {"FLAGS2", 51, 5, 0, 1}, {"REFLECTIVITY2", 56, 8, 0, 1},
    {"SIGNAL", 64, 16, 0, 1}, {"SIGNAL2", 80, 16, 0, 1}, {"ZONE_MASK", 96, 16, 0, 1},
    {"WINDOW", 120, 8, 0, 1}, {"RAW32_WORD1", 0, 32, 0, 1}, {"RAW32_WORD2", 32, 32, 0, 1},
    {"RAW32_WORD3", 64, 32, 0, 1}, {"RAW32_WORD4", 96, 32, 0, 1}, {0, 0, 0, 0, 0}};

/* impl::profiles -- parsing.cpp:327-356 */
static int profile_entry(int profile, const spec_t** tab, size_t* chan_data_size) {
    switch (profile) {
        case ORC_PROFILE_LEGACY: *tab = T_LEGACY; *chan_data_size = 12; return 0;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_DUAL: *tab = T_DUAL; *chan_data_size = 16; return 0;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16: *tab = T_SINGLE; *chan_data_size = 12; return 0;
        case ORC_PROFILE_RNG15_RFL8_NIR8: *tab = T_LB; *chan_data_size = 4; return 0;
        case ORC_PROFILE_FIVE_WORD_PIXEL: *tab = T_FIVE; *chan_data_size = 20; return 0;
        case ORC_PROFILE_FUSA_RNG15_RFL8_NIR8_DUAL:
        case ORC_PROFILE_RNG15_RFL8_NIR8_DUAL: *tab = T_DUAL_LB; *chan_data_size = 8; return 0;
        case ORC_PROFILE_RNG15_RFL8_NIR8_ZONE16: *tab = T_ZM_LB; *chan_data_size = 8; return 0;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_ZONE16:
            *tab = T_ZM_SINGLE; *chan_data_size = 12; return 0;
        case ORC_PROFILE_RNG15_RFL8_WIN8: *tab = T_LB_WIN; *chan_data_size = 4; return 0;
        case ORC_PROFILE_RNG19_RFL8_SIG16_ZONE16_DUAL:
            *tab = T_DUAL_ZONE; *chan_data_size = 16; return 0;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_RGB16: *tab = T_RGB; *chan_data_size = 16; return 0;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_RGB16_DUAL:
            *tab = T_DUAL_RGB; *chan_data_size = 20; return 0;
        default: return -1;
    }
}

/* default LidarFrame field slots -- ouster_core/src/lidar_frame.cpp:73-226 */
typedef struct { const char* name; int ty; } slot_t;
static const slot_t S_LEGACY[] = {{"RANGE", ORC_UINT32}, {"SIGNAL", ORC_UINT16},
    {"NEAR_IR", ORC_UINT16}, {"REFLECTIVITY", ORC_UINT8}, {"FLAGS", ORC_UINT8}, {0, 0}};
static const slot_t S_DUAL[] = {{"RANGE", ORC_UINT32}, {"RANGE2", ORC_UINT32},
    {"SIGNAL", ORC_UINT16}, {"SIGNAL2", ORC_UINT16}, {"REFLECTIVITY", ORC_UINT8},
    {"REFLECTIVITY2", ORC_UINT8}, {"FLAGS", ORC_UINT8}, {"FLAGS2", ORC_UINT8},
    {"NEAR_IR", ORC_UINT16}, {"WINDOW", ORC_UINT8}, {0, 0}};
static const slot_t S_SINGLE[] = {{"RANGE", ORC_UINT32}, {"SIGNAL", ORC_UINT16},
    {"REFLECTIVITY", ORC_UINT8}, {"FLAGS", ORC_UINT8}, {"NEAR_IR", ORC_UINT16},
    {"WINDOW", ORC_UINT8}, {0, 0}};
static const slot_t S_RGB[] = {{"RANGE", ORC_UINT32}, {"SIGNAL", ORC_UINT16},
    {"REFLECTIVITY", ORC_UINT8}, {"NEAR_IR", ORC_UINT16}, {"RGB", ORC_FLOAT16},
    {"FLAGS", ORC_UINT8}, {0, 0}};
static const slot_t S_DUAL_RGB[] = {{"RANGE", ORC_UINT32}, {"RANGE2", ORC_UINT32},
    {"SIGNAL", ORC_UINT16}, {"SIGNAL2", ORC_UINT16}, {"REFLECTIVITY", ORC_UINT8},
    {"REFLECTIVITY2", ORC_UINT8}, {"NEAR_IR", ORC_UINT16}, {"RGB", ORC_FLOAT16},
    {"FLAGS", ORC_UINT8}, {"FLAGS2", ORC_UINT8}, {0, 0}};
static const slot_t S_LB[] = {{"RANGE", ORC_UINT32}, {"REFLECTIVITY", ORC_UINT8},
    {"NEAR_IR", ORC_UINT16}, {"FLAGS", ORC_UINT8}, {0, 0}};
static const slot_t S_LB_WIN[] = {{"RANGE", ORC_UINT32}, {"REFLECTIVITY", ORC_UINT8},
    {"WINDOW", ORC_UINT8}, {"FLAGS", ORC_UINT8}, {0, 0}};
static const slot_t S_ZM_LB[] = {{"RANGE", ORC_UINT32}, {"REFLECTIVITY", ORC_UINT8},
    {"NEAR_IR", ORC_UINT16}, {"FLAGS", ORC_UINT8}, {"ZONE_MASK", ORC_UINT16},
    {"WINDOW", ORC_UINT8}, {0, 0}};
static const slot_t S_ZM_SINGLE[] = {{"RANGE", ORC_UINT32}, {"SIGNAL", ORC_UINT16},
    {"REFLECTIVITY", ORC_UINT8}, {"FLAGS", ORC_UINT8}, {"NEAR_IR", ORC_UINT16},
    {"ZONE_MASK", ORC_UINT16}, {"WINDOW", ORC_UINT8}, {0, 0}};
static const slot_t S_FIVE[] = {{"RAW32_WORD1", ORC_UINT32}, {"RAW32_WORD2", ORC_UINT32},
    {"RAW32_WORD3", ORC_UINT32}, {"RAW32_WORD4", ORC_UINT32}, {"RAW32_WORD5", ORC_UINT32}, {0, 0}};
static const slot_t S_DUAL_LB[] = {{"RANGE", ORC_UINT32}, {"REFLECTIVITY", ORC_UINT8},
    {"NEAR_IR", ORC_UINT16}, {"RANGE2", ORC_UINT32}, {"REFLECTIVITY2", ORC_UINT8},
    {"FLAGS", ORC_UINT8}, {"FLAGS2", ORC_UINT8}, {"WINDOW", ORC_UINT8}, {0, 0}};
static const slot_t S_ZM_DUAL[] = {{"RANGE", ORC_UINT32}, {"RANGE2", ORC_UINT32},
    {"SIGNAL", ORC_UINT16}, {"SIGNAL2", ORC_UINT16}, {"REFLECTIVITY", ORC_UINT8},
    {"REFLECTIVITY2", ORC_UINT8}, {"FLAGS", ORC_UINT8}, {"FLAGS2", ORC_UINT8},
    {"ZONE_MASK", ORC_UINT16}, {"WINDOW", ORC_UINT8}, {0, 0}};

static const slot_t* default_slots(int profile) {
    switch (profile) {
        case ORC_PROFILE_LEGACY: return S_LEGACY;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_DUAL: return S_DUAL;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16: return S_SINGLE;
        case ORC_PROFILE_RNG15_RFL8_NIR8: return S_LB;
        case ORC_PROFILE_RNG15_RFL8_WIN8: return S_LB_WIN;
        case ORC_PROFILE_FIVE_WORD_PIXEL: return S_FIVE;
        case ORC_PROFILE_FUSA_RNG15_RFL8_NIR8_DUAL:
        case ORC_PROFILE_RNG15_RFL8_NIR8_DUAL: return S_DUAL_LB;
        case ORC_PROFILE_RNG15_RFL8_NIR8_ZONE16: return S_ZM_LB;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_ZONE16: return S_ZM_SINGLE;
        case ORC_PROFILE_RNG19_RFL8_SIG16_ZONE16_DUAL: return S_ZM_DUAL;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_RGB16: return S_RGB;
        case ORC_PROFILE_RNG19_RFL8_SIG16_NIR16_RGB16_DUAL: return S_DUAL_RGB;
        default: return NULL;
    }
}

int orc_default_field_type(int profile, const char* name) {
    const slot_t* s = default_slots(profile);
    if (!s) return ORC_VOID;
    for (; s->name; ++s)
        if (strcmp(s->name, name) == 0) return s->ty;
    return ORC_VOID;
}

/* ------------------------------------------------------------------------- */
/* PacketFormat                                                               */
/* ------------------------------------------------------------------------- */

static int cmp_named(const void* a, const void* b) {
    return strcmp(((const orc_named_field*)a)->name, ((const orc_named_field*)b)->name);
}

static orc_field_info fi(size_t bit, size_t size) {
    orc_field_info o;
    orc_field_info_make(bit, size, 0, 0, 1, &o);
    return o;
}

/* PacketFormat::Impl::Impl(const DataFormat&) -- parsing.cpp:453-598 */
int orc_packet_format_init(orc_packet_format* pf, int profile, int header_type,
                           uint32_t pixels_per_column, uint32_t columns_per_packet,
                           uint32_t columns_per_frame) {
    memset(pf, 0, sizeof(*pf));
    const spec_t* tab;
    size_t cds;
    if (profile_entry(profile, &tab, &cds) != 0) return -1;
    int legacy = (profile == ORC_PROFILE_LEGACY);
    int fusa = (header_type == ORC_HEADER_FUSA) && !legacy;

    pf->profile = profile;
    pf->header_type = header_type;
    pf->pixels_per_column = pixels_per_column;
    pf->columns_per_packet = columns_per_packet;
    pf->columns_per_frame = columns_per_frame;
    pf->packet_header_size = legacy ? 0 : 32;
    pf->col_header_size = legacy ? 16 : 12;
    pf->channel_data_size = cds;
    pf->col_footer_size = legacy ? 4 : 0;
    pf->packet_footer_size = legacy ? 0 : 32;
    pf->col_size = pf->col_header_size + pixels_per_column * cds + pf->col_footer_size;
    pf->lidar_packet_size =
        pf->packet_header_size + columns_per_packet * pf->col_size + pf->packet_footer_size;
    if (pf->lidar_packet_size > 65535) return -1;

    for (const spec_t* s = tab; s->name; ++s) {
        orc_named_field* nf = &pf->fields[pf->n_fields++];
        strncpy(nf->name, s->name, ORC_NAME_LEN - 1);
        if (orc_field_info_make(s->bit, s->size, s->up, 0, s->nel, &nf->info) != 0) return -1;
    }
    qsort(pf->fields, (size_t)pf->n_fields, sizeof(orc_named_field), cmp_named);

    /* DataFormat::max_frame_id -- data_format.cpp:163-168 */
    pf->max_frame_id = fusa ? 0xffffffffu : 0xffffu;

    if (legacy) {
        pf->packet_type_info = fi(0, 0);
        pf->init_id_info = fi(0, 0);
        pf->prod_sn_info = fi(0, 0);
        pf->alert_flags_info = fi(0, 0);
        pf->countdown_thermal_shutdown_info = fi(0, 0);
        pf->countdown_shot_limiting_info = fi(0, 0);
        pf->thermal_shutdown_info = fi(0, 0);
        pf->shot_limiting_info = fi(0, 0);
        pf->frame_id_info = fi(80, 16);
        size_t start_bit = 8 * (pf->col_size - pf->col_footer_size);
        if (orc_field_info_make(start_bit, 32, 0, (start_bit + 32) / 8, 1,
                                &pf->col_status_info) != 0)
            return -1;
    } else if (fusa) {
        pf->packet_type_info = fi(0, 8);
        pf->frame_id_info = fi(32, 32);
        pf->init_id_info = fi(8, 24);
        pf->alert_flags_info = fi(64, 8);
        pf->prod_sn_info = fi(88, 40);
        pf->countdown_thermal_shutdown_info = fi(128, 8);
        pf->countdown_shot_limiting_info = fi(136, 8);
        pf->thermal_shutdown_info = fi(144, 4);
        pf->shot_limiting_info = fi(152, 4);
        pf->col_status_info = fi(80, 16);
    } else {
        pf->packet_type_info = fi(0, 16);
        pf->frame_id_info = fi(16, 16);
        pf->init_id_info = fi(32, 24);
        pf->prod_sn_info = fi(56, 40);
        pf->alert_flags_info = fi(96, 8);
        pf->countdown_thermal_shutdown_info = fi(128, 8);
        pf->countdown_shot_limiting_info = fi(136, 8);
        pf->thermal_shutdown_info = fi(144, 4);
        pf->shot_limiting_info = fi(152, 4);
        pf->col_status_info = fi(80, 16);
    }
    pf->col_timestamp_info = fi(0, 64);
    pf->col_measurement_id_info = fi(64, 16);
    return 0;
}

int orc_packet_format_set_fields(orc_packet_format* pf, const orc_named_field* fields, int n,
                                 size_t channel_data_size) {
    if (n > ORC_MAX_FIELDS) return -1;
    memcpy(pf->fields, fields, (size_t)n * sizeof(orc_named_field));
    pf->n_fields = n;
    qsort(pf->fields, (size_t)n, sizeof(orc_named_field), cmp_named);
    pf->channel_data_size = channel_data_size;
    pf->col_size =
        pf->col_header_size + pf->pixels_per_column * channel_data_size + pf->col_footer_size;
    pf->lidar_packet_size =
        pf->packet_header_size + pf->columns_per_packet * pf->col_size + pf->packet_footer_size;
    return 0;
}

const orc_field_info* orc_pf_field(const orc_packet_format* pf, const char* name) {
    for (int i = 0; i < pf->n_fields; ++i)
        if (strcmp(pf->fields[i].name, name) == 0) return &pf->fields[i].info;
    return NULL;
}

/* PacketFormat::block_parsable -- parsing.cpp:958-966 */
int orc_block_parsable(const orc_packet_format* pf) {
    static const int dims[3] = {16, 8, 4};
    for (int i = 0; i < 3; ++i)
        if (pf->pixels_per_column % (uint32_t)dims[i] == 0 &&
            pf->columns_per_packet % (uint32_t)dims[i] == 0)
            return dims[i];
    return 0;
}

/* PacketFormat::frame_id_difference -- parsing.cpp:1312-1321 */
int orc_frame_id_difference(const orc_packet_format* pf, uint32_t current, uint32_t other) {
    int64_t half = pf->max_frame_id >> 1;
    int64_t delta = (int64_t)other - (int64_t)current;
    if (delta < -half) {
        delta += (int64_t)pf->max_frame_id + 1;
    } else if (delta > half) {
        delta -= (int64_t)pf->max_frame_id + 1;
    }
    return (int)delta;
}

/* crc64 (ECMA-182 reflected, Sarwate) -- parsing.cpp:1183-1234 */
uint64_t orc_crc64(const uint8_t* buf, size_t len) {
    static uint64_t table[256];
    static int init = 0;
    if (!init) {
        const uint64_t poly = 0xC96C5795D7870F42ull;
        for (uint32_t i = 0; i < 256; ++i) {
            uint64_t r = i;
            for (int j = 0; j < 8; ++j) r = (r >> 1) ^ (poly & ~((r & 1) - 1));
            table[i] = r;
        }
        init = 1;
    }
    uint64_t crc = ~0ull;
    while (len--) crc = table[(*buf++ ^ (crc & 0xff)) & 0xff] ^ (crc >> 8);
    return ~crc;
}

static inline const uint8_t* nth_col(const orc_packet_format* pf, size_t i, const uint8_t* buf) {
    return buf + pf->packet_header_size + i * pf->col_size; /* parsing.cpp:793-800 */
}

static inline void store_trunc(void* dst, uint64_t word, size_t elem_size) {
    memcpy(dst, &word, elem_size); /* the truncating memcpy of get<T>, field_decode_info.h:51-53 */
}

/* PacketFormat::block_field<T,BlockDim> -- parsing.cpp:628-657 */
int orc_block_field(const orc_packet_format* pf, const char* name, size_t elem_size, void* dst,
                    int cols, const uint8_t* lidar_buf, int block_dim) {
    const orc_field_info* f = orc_pf_field(pf, name);
    if (!f) return -3;
    orc_field_info info = *f;
    if (elem_size < orc_type_size(info.ty_tag) * (size_t)info.num_elements) return -1;
    const uint8_t* col_buf[16];
    uint8_t* data = (uint8_t*)dst;
    for (uint32_t icol = 0; icol < pf->columns_per_packet; icol += (uint32_t)block_dim) {
        for (int i = 0; i < block_dim; ++i) col_buf[i] = nth_col(pf, icol + (uint32_t)i, lidar_buf);
        uint16_t m_id = (uint16_t)orc_field_get(&pf->col_measurement_id_info, col_buf[0]);
        for (uint32_t px = 0; px < pf->pixels_per_column; ++px) {
            ptrdiff_t f_offset = (ptrdiff_t)cols * px + m_id;
            for (int x = 0; x < block_dim; ++x) {
                const uint8_t* px_src =
                    col_buf[x] + pf->col_header_size + px * pf->channel_data_size;
                store_trunc(data + (size_t)(f_offset + x) * elem_size,
                            orc_field_get(&info, px_src), elem_size);
            }
        }
    }
    return 0;
}

/* PacketFormat::col_field<T> -- parsing.cpp:659-675 */
int orc_col_field(const orc_packet_format* pf, const char* name, size_t elem_size,
                  const uint8_t* col_buf, void* dst, int dst_stride) {
    const orc_field_info* f = orc_pf_field(pf, name);
    if (!f) return -3;
    orc_field_info info = *f;
    if (elem_size < orc_type_size(info.ty_tag) * (size_t)info.num_elements) return -1;
    for (uint32_t px = 0; px < pf->pixels_per_column; ++px) {
        const uint8_t* px_src = col_buf + pf->col_header_size + px * pf->channel_data_size;
        store_trunc((uint8_t*)dst + (size_t)px * (size_t)dst_stride * elem_size,
                    orc_field_get(&info, px_src), elem_size);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Frame                                                                      */
/* ------------------------------------------------------------------------- */

int orc_frame_add_field(orc_frame* f, const char* name, int ty_tag) {
    if (f->n_fields >= ORC_MAX_FIELDS) return -1;
    orc_frame_field* ff = &f->fields[f->n_fields];
    memset(ff, 0, sizeof(*ff));
    strncpy(ff->name, name, ORC_NAME_LEN - 1);
    ff->ty_tag = ty_tag;
    ff->elem_size = orc_type_size(ty_tag);
    if (strcmp(name, "RGB") == 0) ff->elem_size *= 3; /* H x W x 3, lidar_frame.cpp:247-252 */
    if (ff->elem_size == 0) return -1;
    ff->data = (uint8_t*)calloc(f->h * f->w, ff->elem_size); /* Field uses calloc, field.cpp:252-254 */
    if (!ff->data) return -1;
    f->n_fields++;
    return 0;
}

/* LidarFrame(w,h,profile,columns_per_packet) -- lidar_frame.cpp:309-360; default fields
 * lidar_frame.cpp:228-256; WINDOW dropped for fw < 3.2 (lidar_frame.cpp:1097-1110) */
orc_frame* orc_frame_create(const orc_packet_format* pf, int with_window) {
    orc_frame* f = (orc_frame*)calloc(1, sizeof(orc_frame));
    if (!f) return NULL;
    f->w = pf->columns_per_frame;
    f->h = pf->pixels_per_column;
    f->n_packets = f->w / pf->columns_per_packet;
    f->frame_id = -1;
    f->timestamp = (uint64_t*)calloc(f->w, 8);
    f->measurement_id = (uint16_t*)calloc(f->w, 2);
    f->status = (uint32_t*)calloc(f->w, 4);
    f->packet_timestamp = (uint64_t*)calloc(f->n_packets ? f->n_packets : 1, 8);
    f->alert_flags = (uint8_t*)calloc(f->n_packets ? f->n_packets : 1, 1);
    const slot_t* s = default_slots(pf->profile);
    for (; s && s->name; ++s) {
        if (!with_window && strcmp(s->name, "WINDOW") == 0) continue;
        orc_frame_add_field(f, s->name, s->ty);
    }
    return f;
}

orc_frame_field* orc_frame_field_by_name(orc_frame* f, const char* name) {
    for (int i = 0; i < f->n_fields; ++i)
        if (strcmp(f->fields[i].name, name) == 0) return &f->fields[i];
    return NULL;
}

void orc_frame_destroy(orc_frame* f) {
    if (!f) return;
    for (int i = 0; i < f->n_fields; ++i) free(f->fields[i].data);
    free(f->timestamp);
    free(f->measurement_id);
    free(f->status);
    free(f->packet_timestamp);
    free(f->alert_flags);
    free(f);
}

/* ------------------------------------------------------------------------- */
/* FrameBatcher                                                               */
/* ------------------------------------------------------------------------- */

typedef struct cached_packet {
    uint8_t* buf;
    size_t len;
    uint64_t ts;
    uint64_t seq;
} cached_packet;

#define ORC_CACHE_CAP 64

struct orc_batcher {
    orc_packet_format pf;
    size_t max_cache_size;
    uint16_t next_valid_m_id;
    cached_packet cache[ORC_CACHE_CAP];
    size_t cache_n;
    uint64_t seq;
    int64_t finished_frame_id, last_frame_id, last_init_id, info_init_id;
    int reset_frame;
    size_t expected_lidar_packets, batched_lidar_packets, dropped_packets;
    int force_col;
};

/* DataFormat::lidar_packets_per_frame -- data_format.cpp:138-161 */
static int lidar_packets_per_frame(const orc_packet_format* pf, uint32_t first, uint32_t second) {
    int start_packet = (int)(first / pf->columns_per_packet);
    int end_packet = (int)(second / pf->columns_per_packet);
    if (second < first) {
        int max_packets = (int)(pf->columns_per_frame / pf->columns_per_packet) +
                          ((pf->columns_per_frame % pf->columns_per_packet) ? 1 : 0);
        int expected = (max_packets - start_packet) + 1 + end_packet;
        if (start_packet == end_packet) return max_packets;
        return expected;
    }
    return end_packet - start_packet + 1;
}

/* FrameBatcher::FrameBatcher -- lidar_frame.cpp:1248-1267 */
orc_batcher* orc_batcher_create(const orc_packet_format* pf, uint32_t init_id, uint32_t cw_first,
                                uint32_t cw_second) {
    if (pf->columns_per_packet == 0 || pf->pixels_per_column == 0) return NULL;
    orc_batcher* b = (orc_batcher*)calloc(1, sizeof(orc_batcher));
    b->pf = *pf;
    b->max_cache_size = 4;
    b->finished_frame_id = -1;
    b->last_frame_id = -1;
    b->last_init_id = init_id;
    b->info_init_id = init_id;
    b->reset_frame = 1;
    b->expected_lidar_packets = (size_t)lidar_packets_per_frame(pf, cw_first, cw_second);
    return b;
}

static void cache_pop_at(orc_batcher* b, size_t i) {
    free(b->cache[i].buf);
    b->cache[i] = b->cache[b->cache_n - 1];
    b->cache_n--;
}

void orc_batcher_destroy(orc_batcher* b) {
    if (!b) return;
    while (b->cache_n) cache_pop_at(b, 0);
    free(b);
}

void orc_batcher_force_col_path(orc_batcher* b, int on) { b->force_col = on; }

static uint32_t pkt_frame_id(const orc_packet_format* pf, const uint8_t* buf) {
    return (uint32_t)orc_field_get(&pf->frame_id_info, buf); /* parsing.cpp:740-742 */
}

/* priority_queue top under PacketComparator (lidar_frame.h:971-993): oldest frame id first */
static size_t cache_top(const orc_batcher* b) {
    size_t best = 0;
    for (size_t i = 1; i < b->cache_n; ++i) {
        int d = orc_frame_id_difference(&b->pf, pkt_frame_id(&b->pf, b->cache[best].buf),
                                        pkt_frame_id(&b->pf, b->cache[i].buf));
        if (d < 0 || (d == 0 && b->cache[i].seq < b->cache[best].seq)) best = i;
    }
    return best;
}

static void cache_packet(orc_batcher* b, const uint8_t* buf, size_t len, uint64_t ts) {
    if (b->cache_n >= ORC_CACHE_CAP) return;
    cached_packet* c = &b->cache[b->cache_n++];
    c->buf = (uint8_t*)malloc(len + 8); /* +8: get<T> may read past the last field */
    memcpy(c->buf, buf, len);
    memset(c->buf + len, 0, 8);
    c->len = len;
    c->ts = ts;
    c->seq = b->seq++;
}

/* zero_header_cols -- lidar_frame.cpp:1274-1278 */
static void zero_header_cols(orc_frame* f, ptrdiff_t start, ptrdiff_t end) {
    if (end <= start) return;
    memset(f->timestamp + start, 0, (size_t)(end - start) * 8);
    memset(f->measurement_id + start, 0, (size_t)(end - start) * 2);
    memset(f->status + start, 0, (size_t)(end - start) * 4);
}

/* zero_field / zero_fields -- lidar_frame.cpp:1371-1418 */
static void zero_fields(orc_frame* f, const orc_packet_format* pf, ptrdiff_t start, ptrdiff_t end) {
    if (start == end) return;
    for (int i = 0; i < pf->n_fields; ++i) {
        orc_frame_field* ff = orc_frame_field_by_name(f, pf->fields[i].name);
        if (!ff) continue;
        size_t row = f->w * ff->elem_size;
        for (size_t u = 0; u < f->h; ++u) {
            uint8_t* p = ff->data + u * row + (size_t)start * ff->elem_size;
            size_t nbytes = (size_t)(end - start) * ff->elem_size;
            if (ff->ty_tag == ORC_FLOAT16) { /* NaN fill 0x7e00, lidar_frame.cpp:1396-1402 */
                for (size_t k = 0; k < nbytes / 2; ++k) {
                    uint16_t v = 0x7e00;
                    memcpy(p + 2 * k, &v, 2);
                }
            } else {
                memset(p, 0, nbytes);
            }
        }
    }
}

/* FrameBatcher::parse_by_col -- lidar_frame.cpp:1422-1466 (RAW_HEADERS not modelled) */
static int parse_by_col(orc_batcher* b, const uint8_t* packet_buf, orc_frame* f) {
    const orc_packet_format* pf = &b->pf;
    for (uint32_t icol = 0; icol < pf->columns_per_packet; icol++) {
        const uint8_t* col_buf = nth_col(pf, icol, packet_buf);
        uint16_t m_id = (uint16_t)orc_field_get(&pf->col_measurement_id_info, col_buf);
        uint64_t ts = orc_field_get(&pf->col_timestamp_info, col_buf);
        uint32_t status = (uint32_t)orc_field_get(&pf->col_status_info, col_buf);
        int valid = (status & 0x01) != 0;
        if (m_id >= f->w) continue;
        if (!valid) continue;
        if (m_id >= b->next_valid_m_id) {
            zero_fields(f, pf, b->next_valid_m_id, m_id);
            zero_header_cols(f, b->next_valid_m_id, m_id);
            b->next_valid_m_id = (uint16_t)(m_id + 1);
        }
        f->timestamp[m_id] = ts;
        f->measurement_id[m_id] = m_id;
        f->status[m_id] = status;
        for (int i = 0; i < pf->n_fields; ++i) { /* foreach_channel_field_ndim */
            orc_frame_field* ff = orc_frame_field_by_name(f, pf->fields[i].name);
            if (!ff) continue;
            int rc = orc_col_field(pf, ff->name, ff->elem_size, col_buf,
                                   ff->data + (size_t)m_id * ff->elem_size, (int)f->w);
            if (rc != 0) return rc;
        }
    }
    return 0;
}

/* FrameBatcher::parse_by_block -- lidar_frame.cpp:1492-1528 */
static int parse_by_block(orc_batcher* b, const uint8_t* packet_buf, orc_frame* f) {
    const orc_packet_format* pf = &b->pf;
    uint16_t first_m_id =
        (uint16_t)orc_field_get(&pf->col_measurement_id_info, nth_col(pf, 0, packet_buf));
    if (first_m_id >= b->next_valid_m_id) {
        zero_fields(f, pf, b->next_valid_m_id, first_m_id);
        zero_header_cols(f, b->next_valid_m_id, first_m_id);
        b->next_valid_m_id = (uint16_t)(first_m_id + pf->columns_per_packet);
    }
    for (uint32_t icol = 0; icol < pf->columns_per_packet; icol++) {
        const uint8_t* col_buf = nth_col(pf, icol, packet_buf);
        uint16_t m_id = (uint16_t)orc_field_get(&pf->col_measurement_id_info, col_buf);
        f->measurement_id[m_id] = m_id;
        f->timestamp[m_id] = orc_field_get(&pf->col_timestamp_info, col_buf);
        f->status[m_id] = (uint32_t)orc_field_get(&pf->col_status_info, col_buf);
    }
    int bd = orc_block_parsable(pf);
    if (bd == 0) return -1;
    for (int i = 0; i < pf->n_fields; ++i) {
        orc_frame_field* ff = orc_frame_field_by_name(f, pf->fields[i].name);
        if (!ff) continue;
        int rc = orc_block_field(pf, ff->name, ff->elem_size, ff->data, (int)f->w, packet_buf, bd);
        if (rc != 0) return rc;
    }
    return 0;
}

/* FrameBatcher::batch_lidar_packet -- lidar_frame.cpp:1530-1576 */
static int batch_lidar_packet(orc_batcher* b, const uint8_t* packet_buf, uint64_t host_ts,
                              orc_frame* f) {
    const orc_packet_format* pf = &b->pf;
    const uint8_t* col0 = nth_col(pf, 0, packet_buf);
    uint16_t packet_id =
        (uint16_t)((uint16_t)orc_field_get(&pf->col_measurement_id_info, col0) /
                   pf->columns_per_packet);
    if (packet_id < f->n_packets) {
        f->packet_timestamp[packet_id] = host_ts;
        f->alert_flags[packet_id] = (uint8_t)orc_field_get(&pf->alert_flags_info, packet_buf);
    }
    size_t block_parsable = (size_t)orc_block_parsable(pf);
    for (uint32_t icol = 0; icol < pf->columns_per_packet; icol++) {
        const uint8_t* col_buf = nth_col(pf, icol, packet_buf);
        uint16_t m_id = (uint16_t)orc_field_get(&pf->col_measurement_id_info, col_buf);
        uint32_t status = (uint32_t)orc_field_get(&pf->col_status_info, col_buf);
        if (!(status & 0x01) || m_id >= f->w) {
            block_parsable = 0;
            break;
        }
    }
    if (block_parsable != 0) {
        for (uint32_t icol = 0; icol < pf->columns_per_packet; icol += (uint32_t)block_parsable) {
            const uint8_t* col_buf = nth_col(pf, icol, packet_buf);
            uint16_t m_id = (uint16_t)orc_field_get(&pf->col_measurement_id_info, col_buf);
            if (m_id + block_parsable > f->w) {
                block_parsable = 0;
                break;
            }
        }
    }
    int rc;
    if (block_parsable != 0 && !b->force_col) {
        rc = parse_by_block(b, packet_buf, f);
    } else {
        rc = parse_by_col(b, packet_buf, f);
    }
    b->batched_lidar_packets++;
    return rc;
}

/* FrameBatcher::start_frame -- lidar_frame.cpp:1709-1741 */
static void start_frame(orc_batcher* b, int64_t f_id, const uint8_t* packet_buf, orc_frame* f) {
    const orc_packet_format* pf = &b->pf;
    b->finished_frame_id = -1;
    b->next_valid_m_id = 0;
    b->batched_lidar_packets = 0;
    f->frame_id = f_id;
    zero_header_cols(f, 0, (ptrdiff_t)f->w);
    memset(f->packet_timestamp, 0, f->n_packets * 8);
    uint8_t therm = (uint8_t)orc_field_get(&pf->thermal_shutdown_info, packet_buf);
    uint8_t shot = (uint8_t)orc_field_get(&pf->shot_limiting_info, packet_buf);
    /* frame_status() -- lidar_frame.cpp:1310-1323 */
    f->frame_status = (uint64_t)((therm & 0x0f) << 0) | (uint64_t)((shot & 0x0f) << 4);
    f->shutdown_countdown =
        (uint8_t)(uint16_t)orc_field_get(&pf->countdown_thermal_shutdown_info, packet_buf);
    f->shot_limiting_countdown =
        (uint8_t)(uint16_t)orc_field_get(&pf->countdown_shot_limiting_info, packet_buf);
}

/* FrameBatcher::check_frame_complete -- lidar_frame.cpp:1894-1903 */
static int check_frame_complete(const orc_batcher* b, const orc_frame* f) {
    size_t nz = 0;
    for (size_t i = 0; i < f->n_packets; ++i) nz += (f->packet_timestamp[i] != 0);
    return b->batched_lidar_packets >= b->expected_lidar_packets &&
           nz == b->expected_lidar_packets;
}

/* FrameBatcher::finalize_frame -- lidar_frame.cpp:1905-1927; returns -2 for the FUSA throw */
static int finalize_frame(orc_batcher* b, orc_frame* f) {
    if (b->next_valid_m_id < f->w) zero_fields(f, &b->pf, b->next_valid_m_id, (ptrdiff_t)f->w);
    if (b->info_init_id == b->last_init_id && f->frame_id <= b->last_frame_id &&
        b->pf.header_type == ORC_HEADER_FUSA)
        return -2;
    b->finished_frame_id = f->frame_id;
    b->last_frame_id = f->frame_id;
    b->batched_lidar_packets = 0;
    return 1;
}

/* FrameBatcher::reset -- lidar_frame.cpp:1929-1940 */
void orc_batcher_reset(orc_batcher* b) {
    b->reset_frame = 1;
    b->finished_frame_id = -1;
    b->next_valid_m_id = 0;
    b->batched_lidar_packets = 0;
    while (b->cache_n) cache_pop_at(b, 0);
}

/* FrameBatcher::batch_with_caching -- lidar_frame.cpp:1743-1793 */
static int batch_with_caching(orc_batcher* b, const uint8_t* buf, size_t len, uint64_t ts,
                              orc_frame* f) {
    cache_packet(b, buf, len, ts);
    while (b->cache_n) {
        size_t top = cache_top(b);
        const uint8_t* pbuf = b->cache[top].buf;
        int64_t f_id = pkt_frame_id(&b->pf, pbuf);
        if (b->finished_frame_id >= 0 &&
            orc_frame_id_difference(&b->pf, (uint32_t)b->finished_frame_id, (uint32_t)f_id) <= 0) {
            b->dropped_packets++;
            cache_pop_at(b, top);
            continue;
        }
        if (f->frame_id == -1 || b->finished_frame_id >= 0) start_frame(b, f_id, pbuf, f);
        int diff = orc_frame_id_difference(&b->pf, (uint32_t)f->frame_id, (uint32_t)f_id);
        if (diff < 0) {
            b->dropped_packets++;
            cache_pop_at(b, top);
        } else if (diff > 0) {
            if (b->cache_n >= b->max_cache_size) return finalize_frame(b, f);
            return 0;
        } else {
            int rc = batch_lidar_packet(b, pbuf, b->cache[top].ts, f);
            cache_pop_at(b, top);
            if (rc != 0) return rc;
            if (check_frame_complete(b, f)) return finalize_frame(b, f);
        }
    }
    return 0;
}

/* FrameBatcher::handle_init_id_change -- lidar_frame.cpp:1795-1822 */
static int handle_init_id_change(orc_batcher* b, const uint8_t* buf, size_t len, uint64_t ts,
                                 orc_frame* f) {
    b->last_init_id = (int64_t)(uint32_t)orc_field_get(&b->pf.init_id_info, buf);
    if (f->frame_id == -1 || b->finished_frame_id >= 0) {
        orc_batcher_reset(b);
        b->reset_frame = 0;
        int64_t f_id = pkt_frame_id(&b->pf, buf);
        start_frame(b, f_id, buf, f);
        int rc = batch_lidar_packet(b, buf, ts, f);
        if (rc != 0) return rc;
        if (check_frame_complete(b, f)) return finalize_frame(b, f);
        return 0;
    }
    int rc = finalize_frame(b, f);
    if (rc < 0) return rc;
    orc_batcher_reset(b);
    cache_packet(b, buf, len, ts);
    return 1;
}

/* FrameBatcher::batch -- lidar_frame.cpp:1824-1884 (lidar packets only) */
int orc_batcher_batch(orc_batcher* b, const uint8_t* buf, size_t len, uint64_t host_timestamp,
                      orc_frame* f) {
    if (b->reset_frame) {
        f->frame_id = -1;
        b->reset_frame = 0;
    }
    if (f->w != b->pf.columns_per_frame || f->h != b->pf.pixels_per_column) return -1;
    if (f->n_packets != f->w / b->pf.columns_per_packet) return -1;

    /* work on a padded private copy: FieldDecodeInfo::get reads 8 bytes per field */
    uint8_t* pbuf = (uint8_t*)malloc(len + 8);
    memcpy(pbuf, buf, len);
    memset(pbuf + len, 0, 8);
    int ret = 0;

    if (b->pf.profile != ORC_PROFILE_LEGACY &&
        (int64_t)(uint32_t)orc_field_get(&b->pf.init_id_info, pbuf) != b->last_init_id) {
        ret = handle_init_id_change(b, pbuf, len, host_timestamp, f);
        free(pbuf);
        return ret;
    }
    int64_t f_id = pkt_frame_id(&b->pf, pbuf);
    if (b->cache_n == 0) {
        if (b->finished_frame_id >= 0 &&
            orc_frame_id_difference(&b->pf, (uint32_t)b->finished_frame_id, (uint32_t)f_id) <= 0) {
            b->dropped_packets++;
            free(pbuf);
            return 0;
        }
        if (f->frame_id == -1 || b->finished_frame_id >= 0) {
            start_frame(b, f_id, pbuf, f);
            ret = batch_lidar_packet(b, pbuf, host_timestamp, f);
            if (ret == 0 && check_frame_complete(b, f)) ret = finalize_frame(b, f);
            free(pbuf);
            return ret;
        }
    }
    if (f->frame_id == f_id && b->finished_frame_id < 0) {
        ret = batch_lidar_packet(b, pbuf, host_timestamp, f);
        if (ret == 0 && check_frame_complete(b, f)) ret = finalize_frame(b, f);
        free(pbuf);
        return ret;
    }
    ret = batch_with_caching(b, pbuf, len, host_timestamp, f);
    free(pbuf);
    return ret;
}

size_t orc_batcher_batched_packets(const orc_batcher* b) { return b->batched_lidar_packets; }
size_t orc_batcher_dropped_packets(const orc_batcher* b) { return b->dropped_packets; }
int orc_batcher_set_max_cache_size(orc_batcher* b, size_t n) {
    if (n == 0) return -1;
    b->max_cache_size = n;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* frame_to_packets (lidar part) -- impl/lidar_frame_impl.h:435-531           */
/* PacketFormat::set_block -- parsing.cpp:1056-1090                           */
/* ------------------------------------------------------------------------- */
int orc_frame_to_packets(const orc_frame* f, const orc_packet_format* pf, uint32_t init_id,
                         uint64_t prod_sn, uint8_t* out, uint64_t* ts_out) {
    size_t total = f->n_packets;
    if (f->w / pf->columns_per_packet != total) return -1;
    size_t psz = pf->lidar_packet_size;
    int emitted = 0;
    uint8_t* buf = (uint8_t*)malloc(psz + 8);
    for (size_t packet_id = 0; packet_id < total; ++packet_id) {
        memset(buf, 0, psz + 8);
        uint64_t host_ts = f->packet_timestamp[packet_id];
        /* set_header lambda, impl/lidar_frame_impl.h:457-470 */
        orc_field_set(&pf->thermal_shutdown_info, buf, f->frame_status & 0x0f);
        orc_field_set(&pf->shot_limiting_info, buf, (f->frame_status & 0xf0) >> 4);
        orc_field_set(&pf->countdown_thermal_shutdown_info, buf, f->shutdown_countdown);
        orc_field_set(&pf->countdown_shot_limiting_info, buf, f->shot_limiting_countdown);
        orc_field_set(&pf->frame_id_info, buf, (uint32_t)f->frame_id);
        orc_field_set(&pf->init_id_info, buf, init_id);
        orc_field_set(&pf->prod_sn_info, buf, prod_sn);
        orc_field_set(&pf->packet_type_info, buf, 0x1);
        orc_field_set(&pf->alert_flags_info, buf, f->alert_flags[packet_id]);

        int any_valid = 0;
        int valid[64];
        uint32_t cpp = pf->columns_per_packet;
        for (uint32_t icol = 0; icol < cpp; ++icol) {
            uint8_t* col_buf = (uint8_t*)nth_col(pf, icol, buf);
            size_t id = packet_id * cpp + icol;
            orc_field_set(&pf->col_status_info, col_buf, f->status[id]);
            orc_field_set(&pf->col_measurement_id_info, col_buf, (uint16_t)id);
            orc_field_set(&pf->col_timestamp_info, col_buf, f->timestamp[id]);
            any_valid |= (int)(f->status[id] & 0x01);
        }
        if (!any_valid && !host_ts) continue;

        /* set_block per frame field that the profile carries */
        for (uint32_t i = 0; i < cpp; ++i)
            valid[i] = (int)(orc_field_get(&pf->col_status_info, nth_col(pf, i, buf)) & 0x01);
        uint16_t m_id0 =
            (uint16_t)orc_field_get(&pf->col_measurement_id_info, nth_col(pf, 0, buf));
        for (int fi_ = 0; fi_ < pf->n_fields; ++fi_) {
            const orc_frame_field* ff = NULL;
            for (int k = 0; k < f->n_fields; ++k)
                if (strcmp(f->fields[k].name, pf->fields[fi_].name) == 0) ff = &f->fields[k];
            if (!ff) continue;
            const orc_field_info* info = &pf->fields[fi_].info;
            for (uint32_t px = 0; px < pf->pixels_per_column; ++px) {
                size_t f_offset = f->w * px + m_id0;
                for (uint32_t x = 0; x < cpp; ++x) {
                    if (!valid[x]) continue;
                    uint8_t* px_dst = (uint8_t*)nth_col(pf, x, buf) + pf->col_header_size +
                                      px * pf->channel_data_size;
                    uint64_t v = 0;
                    memcpy(&v, ff->data + (f_offset + x) * ff->elem_size, ff->elem_size);
                    orc_field_set(info, px_dst, v);
                }
            }
        }
        if (pf->profile != ORC_PROFILE_LEGACY && pf->header_type == ORC_HEADER_STANDARD) {
            uint64_t crc = orc_crc64(buf, psz - 8);
            memcpy(buf + psz - 8, &crc, 8);
        }
        memcpy(out + (size_t)emitted * psz, buf, psz);
        if (ts_out) ts_out[emitted] = host_ts;
        emitted++;
    }
    free(buf);
    return emitted;
}

/* ------------------------------------------------------------------------- */
/* destagger -- impl/lidar_frame_impl.h:733-760 (2-D) and :776-811 (N-D, k>1)  */
/* returns -1: "image height does not match shifts size"                       */
/* ------------------------------------------------------------------------- */
int orc_destagger(size_t elem_size, size_t k, const void* img, const int* shifts, size_t n_shifts,
                  size_t h, size_t w, int inverse, void* out) {
    if (n_shifts != h) return -1;
    int sign = inverse ? -1 : +1;
    const uint8_t* g = (const uint8_t*)img;
    uint8_t* d = (uint8_t*)out;
    size_t px = elem_size * k; /* bytes per pixel incl. trailing dims */
    for (size_t u = 0; u < h; ++u) {
        const uint8_t* g_row = g + u * w * px;
        uint8_t* d_row = d + u * w * px;
        /* literal restatement of `(w + sign * shift[u] % w) % w` with w a size_t: the int
         * product is converted to uint64 before `%` (impl/lidar_frame_impl.h:756) */
        const int offset = (int)((w + (size_t)(sign * shifts[u]) % w) % w);
        memcpy(d_row, g_row + (w - (size_t)offset) * px, (size_t)offset * px);
        memcpy(d_row + (size_t)offset * px, g_row, (w - (size_t)offset) * px);
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* cartesianT<T> -- impl/cartesian.h:36-66                                     */
/* ------------------------------------------------------------------------- */
#define ORC_CART_BODY(T)                                         \
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {               \
        const uint32_t r = rng[i];                               \
        const ptrdiff_t ix = i * 3, iy = i * 3 + 1, iz = i * 3 + 2; \
        if (r == 0) {                                            \
            pts[ix] = pts[iy] = pts[iz] = (T)0.0;                \
        } else {                                                 \
            pts[ix] = r * dir[ix] + ofs[ix];                     \
            pts[iy] = r * dir[iy] + ofs[iy];                     \
            pts[iz] = r * dir[iz] + ofs[iz];                     \
        }                                                        \
    }

void orc_cartesian_f64(double* pts, const uint32_t* rng, const double* dir, const double* ofs,
                       size_t n) {
    ORC_CART_BODY(double)
}
void orc_cartesian_f32(float* pts, const uint32_t* rng, const float* dir, const float* ofs,
                       size_t n) {
    ORC_CART_BODY(float)
}
void orc_cartesian_f64_omp(double* pts, const uint32_t* rng, const double* dir, const double* ofs,
                           size_t n) {
#pragma omp parallel for schedule(static)
    ORC_CART_BODY(double)
}
void orc_cartesian_f32_omp(float* pts, const uint32_t* rng, const float* dir, const float* ofs,
                           size_t n) {
#pragma omp parallel for schedule(static)
    ORC_CART_BODY(float)
}

/* ------------------------------------------------------------------------- */
/* make_xyz_lut -- ouster_core/src/xyzlut.cpp:11-89                            */
/* returns -1 "lut dimensions must be greater than zero", -2 "unexpected frame dimensions" */
/* ------------------------------------------------------------------------- */
int orc_make_xyz_lut(size_t w, size_t h, double range_unit, const double* b2l, const double* tr,
                     const double* az_deg, size_t n_az, const double* alt_deg, size_t n_alt,
                     double* direction, double* offset) {
    if (w == 0 || h == 0) return -1;
    if ((n_az != h || n_alt != h) && (n_az != w * h || n_alt != w * h)) return -2;

    double b03 = b2l[0 * 4 + 3], b23 = b2l[2 * 4 + 3];
    double dist = b03;
    if (b23 != 0) dist = sqrt(pow(b03, 2) + pow(b23, 2));

    int per_beam = (n_az == h && n_alt == h);
    const double azimuth_radians = M_PI * 2.0 / (double)w;
    for (size_t row = 0; row < h; ++row) {
        for (size_t col = 0; col < w; ++col) {
            size_t i = row * w + col;
            double enc, az, alt;
            if (per_beam) {
                enc = 2.0 * M_PI - ((double)col * azimuth_radians);
                az = -az_deg[row] * M_PI / 180.0;
                alt = alt_deg[row] * M_PI / 180.0;
            } else {
                enc = 0;
                az = az_deg[i] * M_PI / 180.0;
                alt = alt_deg[i] * M_PI / 180.0;
            }
            double d[3], o[3];
            d[0] = cos(enc + az) * cos(alt);
            d[1] = sin(enc + az) * cos(alt);
            d[2] = sin(alt);
            o[0] = cos(enc) * b03 - d[0] * dist;
            o[1] = sin(enc) * b03 - d[1] * dist;
            o[2] = -d[2] * dist + b23;
            /* row-vector * R^T, then + t  (xyzlut.cpp:78-82) */
            for (int j = 0; j < 3; ++j) {
                double dj = d[0] * tr[j * 4 + 0] + d[1] * tr[j * 4 + 1] + d[2] * tr[j * 4 + 2];
                double oj = o[0] * tr[j * 4 + 0] + o[1] * tr[j * 4 + 1] + o[2] * tr[j * 4 + 2];
                oj += tr[j * 4 + 3];
                direction[i * 3 + (size_t)j] = dj * range_unit;
                offset[i * 3 + (size_t)j] = oj * range_unit;
            }
        }
    }
    return 0;
}

/* dewarp<T> -- pose_util.h:37-59 (transform<T> is the W == 1 case, :118-131) */
#define ORC_DEWARP_BODY(T)                                                              \
    for (size_t wv = 0; wv < w; ++wv) {                                                 \
        const T* m = poses + wv * 16;                                                   \
        const size_t hh = n / w;                                                        \
        for (size_t i = 0; i < hh; ++i) {                                               \
            const size_t ix = i * w + wv;                                               \
            const T x = pts[ix * 3], y = pts[ix * 3 + 1], z = pts[ix * 3 + 2];         \
            for (int r = 0; r < 3; ++r) {                                               \
                const T a = m[r * 4] * x, b = m[r * 4 + 1] * y, c = m[r * 4 + 2] * z;  \
                out[ix * 3 + r] = (a + (b + c)) + m[r * 4 + 3];                         \
            }                                                                           \
        }                                                                               \
    }

void orc_dewarp_f64(double* out, const double* pts, const double* poses, size_t n, size_t w) {
    ORC_DEWARP_BODY(double)
}
void orc_dewarp_f32(float* out, const float* pts, const float* poses, size_t n, size_t w) {
    ORC_DEWARP_BODY(float)
}

/* dewarp(LidarFrame, XYZLutT<T>, min_range, max_range) -- impl/dewarp_impl.h:22-76 (single frame):
 * cartesian of the whole range image, then, column by column between the first and last column
 * whose status has bit 0 set (lidar_frame.cpp:907-925), skipping columns whose status word is 0,
 * every pixel with min_r <= r <= max_r is posed with the column's body_to_world (cast to T) and
 * appended; optional per-point column index and column timestamp.  Returns the point count. */
#define ORC_DEWARP_FRAME_BODY(T, CART)                                                    \
    const size_t n = h * w;                                                               \
    T* pts = (T*)malloc(n * 3 * sizeof(T));                                               \
    if (!pts) return 0;                                                                   \
    CART(pts, range, dir, off, n);                                                        \
    const uint32_t min_r = (uint32_t)ceil(min_range * 1e3);                               \
    const uint32_t max_r = (uint32_t)floor(max_range * 1e3);                              \
    long start_col = -1, stop_col = -1;                                                   \
    for (size_t i = 0; i < w; ++i)                                                        \
        if ((status[i] & 1u) > 0) {                                                       \
            start_col = (long)i;                                                          \
            break;                                                                        \
        }                                                                                 \
    for (long i = (long)w - 1; i >= 0; --i)                                               \
        if ((status[i] & 1u) > 0) {                                                       \
            stop_col = i;                                                                 \
            break;                                                                        \
        }                                                                                 \
    size_t count = 0;                                                                     \
    if (start_col >= 0 && stop_col >= start_col) {                                        \
        for (long x = start_col; x <= stop_col; ++x) {                                    \
            if (status[x] == 0) continue;                                                 \
            T m[12];                                                                      \
            for (int k = 0; k < 12; ++k) m[k] = (T)poses[(size_t)x * 16 + (size_t)k];     \
            for (size_t y = 0; y < h; ++y) {                                              \
                const uint32_t r = range[y * w + (size_t)x];                              \
                if (r >= min_r && r <= max_r) {                                           \
                    const T* p = pts + (y * w + (size_t)x) * 3;                           \
                    for (int c = 0; c < 3; ++c) {                                         \
                        const T a = m[c * 4] * p[0], b = m[c * 4 + 1] * p[1], cc = m[c * 4 + 2] * p[2]; \
                        out[count * 3 + (size_t)c] = (a + (b + cc)) + m[c * 4 + 3];       \
                    }                                                                     \
                    if (col_idx) col_idx[count] = (uint32_t)x;                            \
                    if (ts_out) ts_out[count] = timestamps[x];                            \
                    ++count;                                                              \
                }                                                                         \
            }                                                                             \
        }                                                                                 \
    }                                                                                     \
    free(pts);                                                                            \
    return count;

size_t orc_dewarp_frame_f64(double* out, uint32_t* col_idx, uint64_t* ts_out, const uint32_t* range,
                            const double* dir, const double* off, const double* poses,
                            const uint32_t* status, const uint64_t* timestamps, size_t h, size_t w,
                            double min_range, double max_range) {
    ORC_DEWARP_FRAME_BODY(double, orc_cartesian_f64)
}
size_t orc_dewarp_frame_f32(float* out, uint32_t* col_idx, uint64_t* ts_out, const uint32_t* range,
                            const float* dir, const float* off, const double* poses,
                            const uint32_t* status, const uint64_t* timestamps, size_t h, size_t w,
                            double min_range, double max_range) {
    ORC_DEWARP_FRAME_BODY(float, orc_cartesian_f32)
}

/* matrix_hash -- tests/frame_batcher_test.cpp:595-606 (libstdc++ std::hash<int> = identity) */
uint64_t orc_snapshot_hash(const void* data, size_t n, size_t elem_size) {
    uint64_t seed = 0;
    const uint8_t* p = (const uint8_t*)data;
    for (size_t i = 0; i < n; ++i) {
        uint64_t e = 0;
        memcpy(&e, p + i * elem_size, elem_size);
        seed ^= e + 0x9e3779b9ull + (seed << 6) + (seed >> 2);
    }
    return seed;
}
