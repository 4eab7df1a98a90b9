/*
 * rtw_oracle_impl.h -- type-generic body of the CPU oracle (TEST INFRASTRUCTURE, see rtw_oracle.h).
 * Included twice by rtw_oracle.c with
 *     RT      float | double                  element type T of the reference's generic code
 *     SFX(x)  x##_f32 | x##_f64
 *     FMA/SQRT/FABS/FMIN  the matching libm entry points
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 */

typedef struct { RT x, y, z; } SFX(v3);

static inline SFX(v3) SFX(mk)(RT x, RT y, RT z) { SFX(v3) r = {x, y, z}; return r; }

/* StaticArrays dot, contracted under @fastmath: fma(z,z, fma(y,y, x*x)) -- see FP contract in rtw_oracle.h */
static inline RT SFX(dot)(SFX(v3) a, SFX(v3) b) { return FMA(a.z, b.z, FMA(a.y, b.y, a.x * b.x)); }

/* src/vec.jl:19 */
static inline RT SFX(squared_length)(SFX(v3) a) { return SFX(dot)(a, a); }

/* src/vec.jl:20 -- the literal 1e-5 is Float64, so a Float32 squared length is promoted before comparing */
static inline int SFX(near_zero)(SFX(v3) a) { return (double)SFX(squared_length)(a) < 1e-5; }

/* StaticArrays 1.2.13 normalize(v) = inv(norm(v)) * v */
static inline SFX(v3) SFX(normalize)(SFX(v3) a) {
    RT inv = (RT)1 / SQRT(SFX(dot)(a, a));
    return SFX(mk)(a.x * inv, a.y * inv, a.z * inv);
}

/* ---------------------------------------------------------------- RNG draws, src/rand.jl */

/* trand(T), src/rand.jl:10-13.  Float32: 23 random mantissa bits; Float64: 52. */
static inline RT SFX(trand)(rtwo_rng* g) {
#if RT_IS_F32
    return (RT)(rtwo_next_u32(g) >> 9) * (RT)(1.0 / 8388608.0);
#else
    return (RT)(rtwo_next_u64(g) >> 12) * (RT)(1.0 / 4503599627370496.0);
#endif
}

/* random_between(min,max) = trand(T)*(max-min) + min, src/rand.jl:24 (contracted to one fma) */
static inline RT SFX(random_between)(rtwo_rng* g, RT lo, RT hi) { return FMA(SFX(trand)(g), hi - lo, lo); }

/* random_vec3_in_sphere, src/rand.jl:15-22: rejection in [-1,1]^3, x,y,z drawn in order.
 * Production stream: attempt a uses draws 4a..4a+2 of the current scatter event. */
static inline SFX(v3) SFX(random_vec3_in_sphere)(rtwo_rng* g) {
    for (uint32_t a = 0;; ++a) {
        SFX(v3) p;
        rtwo_rng_seek(g, 4u * a, RT_WORDS_PER_DRAW);
        p.x = SFX(random_between)(g, (RT)-1, (RT)1);
        p.y = SFX(random_between)(g, (RT)-1, (RT)1);
        p.z = SFX(random_between)(g, (RT)-1, (RT)1);
        if (SFX(dot)(p, p) <= (RT)1) return p;
    }
}

/* random_vec3_on_sphere, src/rand.jl:29 */
static inline SFX(v3) SFX(random_vec3_on_sphere)(rtwo_rng* g) {
    return SFX(normalize)(SFX(random_vec3_in_sphere)(g));
}

/* random_vec2_in_disk, src/rand.jl:31-38 (2-vector dot = fma(y,y, x*x)) */
static inline void SFX(random_vec2_in_disk)(rtwo_rng* g, RT* px, RT* py) {
    rtwo_rng_seek(g, 2u, RT_WORDS_PER_DRAW); /* production stream: disk attempt k = draws 2+2k, 3+2k of event 0 */
    for (;;) {
        RT x = SFX(random_between)(g, (RT)-1, (RT)1);
        RT y = SFX(random_between)(g, (RT)-1, (RT)1);
        if (FMA(y, y, x * x) <= (RT)1) { *px = x; *py = y; return; }
    }
}

/* ---------------------------------------------------------------- light transport, src/light.jl */

/* reflect(v,n) = v - (2v.n)*n, src/light.jl:6.  (2v).n == 2*(v.n) exactly (power-of-two scaling). */
static inline SFX(v3) SFX(reflect)(SFX(v3) v, SFX(v3) n) {
    RT k = (RT)2 * SFX(dot)(v, n);
    return SFX(mk)(FMA(-k, n.x, v.x), FMA(-k, n.y, v.y), FMA(-k, n.z, v.z));
}

/* refract(dir,n,ratio), src/light.jl:12-17 */
static inline SFX(v3) SFX(refract)(SFX(v3) d, SFX(v3) n, RT ratio) {
    RT cos_t = FMIN(-SFX(dot)(d, n), (RT)1);
    SFX(v3) perp = SFX(mk)(ratio * FMA(cos_t, n.x, d.x), ratio * FMA(cos_t, n.y, d.y), ratio * FMA(cos_t, n.z, d.z));
    RT s = SQRT(FABS((RT)1 - SFX(squared_length)(perp)));
    /* r_out_perp + (-s)*n */
    return SFX(normalize)(SFX(mk)(FMA(-s, n.x, perp.x), FMA(-s, n.y, perp.y), FMA(-s, n.z, perp.z)));
}

/* reflectance(cos,ratio): Schlick, src/light.jl:19-25 */
static inline RT SFX(reflectance)(RT cos_t, RT ratio) {
    RT r0 = ((RT)1 - ratio) / ((RT)1 + ratio);
    r0 = r0 * r0;
    RT x = (RT)1 - cos_t;
    RT x2 = x * x;
    RT x4 = x2 * x2;
    RT x5 = x4 * x;
    return FMA((RT)1 - r0, x5, r0);
}

/* ---------------------------------------------------------------- intersection, src/hit.jl */

typedef struct {
    RT t;
    SFX(v3) p;
    SFX(v3) n;
    int front_face;
    uint32_t index; /* which sphere: stands in for the boxed `mat` field of HitRecord, src/structs.jl:26 */
} SFX(hitrec);

/* hit(s::Sphere, r, tmin, tmax), src/hit.jl:12-35 + ray_to_HitRecord src/hit.jl:6-10 + point src/hit.jl:3 */
static inline int SFX(hit_sphere)(SFX(v3) c, RT radius, SFX(v3) o, SFX(v3) d, RT tmin, RT tmax, SFX(hitrec)* rec) {
    SFX(v3) oc = SFX(mk)(o.x - c.x, o.y - c.y, o.z - c.z);  /* :13 */
    RT half_b = SFX(dot)(oc, d);                            /* :16 (a = 1, :15) */
    RT cq = FMA(-radius, radius, SFX(dot)(oc, oc));         /* :17 oc.oc - radius^2 */
    RT disc = FMA(half_b, half_b, -cq);                     /* :18 half_b^2 - a*c */
    if (disc < (RT)0) return 0;                             /* :19 (a NaN discriminant is NOT a miss, as in the reference) */
    RT sqrtd = SQRT(disc);                                  /* :20 */
    RT root = -half_b - sqrtd;                              /* :23 */
    if (root < tmin || tmax < root) {                       /* :24 */
        root = -half_b + sqrtd;                             /* :25 */
        if (root < tmin || tmax < root) return 0;           /* :26-28 */
    }
    rec->t = root;
    rec->p = SFX(mk)(FMA(root, d.x, o.x), FMA(root, d.y, o.y), FMA(root, d.z, o.z)); /* :32, point :3 */
    SFX(v3) on = SFX(mk)((rec->p.x - c.x) / radius, (rec->p.y - c.y) / radius, (rec->p.z - c.z) / radius); /* :33 */
    rec->front_face = SFX(dot)(d, on) < (RT)0;              /* :7 */
    rec->n = rec->front_face ? on : SFX(mk)(-on.x, -on.y, -on.z); /* :8 */
    return 1;
}

typedef struct {
    const RT* geom4;
    const RT* mat4;
    const uint32_t* kind;
    uint32_t n;
    uint64_t segments;
    uint64_t tests;
    /* optional debugging record of a single path (rtwo_path_trace_*): 8 doubles per segment -- origin, direction, index
     * of the closest sphere (-1: miss), its t */
    double* trace;
    int trace_cap, trace_n;
} SFX(world);

/* hit(hittables::HittableList, r, tmin, tmax), src/hit.jl:38-50 */
static inline int SFX(hit_list)(SFX(world)* w, SFX(v3) o, SFX(v3) d, RT tmin, RT tmax, SFX(hitrec)* best) {
    RT closest = tmax; /* :39 */
    int any = 0;       /* :40 */
    SFX(hitrec) rec;
    for (uint32_t i = 0; i < w->n; ++i) { /* :41 */
        const RT* g = w->geom4 + 4 * (size_t)i;
        if (SFX(hit_sphere)(SFX(mk)(g[0], g[1], g[2]), g[3], o, d, tmin, closest, &rec)) { /* :43-44 */
            rec.index = i;
            *best = rec;      /* :45 */
            closest = rec.t;  /* :46 */
            any = 1;
        }
    }
    w->segments += 1;
    w->tests += w->n;
    if (w->trace && w->trace_n < w->trace_cap) {
        double* r = w->trace + 8 * (size_t)w->trace_n++;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
        r[6] = any ? (double)best->index : -1.0;
        r[7] = any ? (double)best->t : 0.0;
    }
    return any;
}

/* ---------------------------------------------------------------- materials, src/material.jl */

/* scatter(): returns the scattered direction; *att = attenuation.  Origin of the new ray is rec.p. */
static inline SFX(v3) SFX(scatter)(SFX(world)* w, rtwo_rng* g, SFX(v3) d_in, const SFX(hitrec)* rec, SFX(v3)* att) {
    const RT* m = w->mat4 + 4 * (size_t)rec->index;
    uint32_t kind = w->kind[rec->index];
    if (kind == RTWO_LAMBERTIAN) { /* src/material.jl:13-23 */
        SFX(v3) rv = SFX(random_vec3_on_sphere)(g);
        SFX(v3) sd = SFX(mk)(rec->n.x + rv.x, rec->n.y + rv.y, rec->n.z + rv.z); /* :14 */
        if (SFX(near_zero)(sd)) sd = rec->n;                                       /* :15-16 */
        else sd = SFX(normalize)(sd);                                              /* :18 */
        *att = SFX(mk)(m[0], m[1], m[2]);                                          /* :21 */
        return sd;
    } else if (kind == RTWO_METAL) { /* src/material.jl:31-34; the unit vector is drawn even when fuzz == 0 */
        SFX(v3) refl = SFX(reflect)(d_in, rec->n);
        SFX(v3) rv = SFX(random_vec3_on_sphere)(g);
        RT fuzz = m[3];
        *att = SFX(mk)(m[0], m[1], m[2]);
        return SFX(normalize)(SFX(mk)(FMA(fuzz, rv.x, refl.x), FMA(fuzz, rv.y, refl.y), FMA(fuzz, rv.z, refl.z)));
    } else { /* Dielectric, src/material.jl:41-53 */
        RT ir = m[3];
        RT ratio = rec->front_face ? ((RT)1 / ir) : ir;       /* :43 */
        RT cos_t = FMIN(-SFX(dot)(d_in, rec->n), (RT)1);      /* :44 */
        RT sin_t = SQRT(FMA(-cos_t, cos_t, (RT)1));           /* :45 */
        int cannot_refract = ratio * sin_t > (RT)1;           /* :46 */
        *att = SFX(mk)((RT)1, (RT)1, (RT)1);                  /* :42 */
        /* :47 `||` short-circuits: no RNG draw on total internal reflection (production stream: draw 3 of the event) */
        rtwo_rng_seek(g, 3u, RT_WORDS_PER_DRAW);
        if (cannot_refract || SFX(reflectance)(cos_t, ratio) > SFX(trand)(g))
            return SFX(reflect)(d_in, rec->n);                /* :48 (not re-normalised) */
        return SFX(refract)(d_in, rec->n, ratio);             /* :50 */
    }
}

/* ---------------------------------------------------------------- integrator, src/ray_color.jl */

/* skycolor(ray), src/ray_color.jl:1-6: t in T, constants in Float64, not @fastmath (no contraction) */
static inline void SFX(skycolor)(SFX(v3) d, double out[3]) {
    RT t = (RT)0.5 * (d.y + (RT)1);
    double a = (double)((RT)1 - t), b = (double)t;
    out[0] = a * 1.0 + b * 0.5;
    out[1] = a * 1.0 + b * 0.7;
    out[2] = a * 1.0 + b * 1.0;
}

/* ray_color(r, world, depth), src/ray_color.jl:14-38 -- recursive exactly like the reference;
 * attenuation (T) .* colour (Float64) multiplies innermost-first (:31). */
static void SFX(ray_color)(SFX(world)* w, rtwo_rng* g, SFX(v3) o, SFX(v3) d, int depth, double out[3]) {
    if (depth <= 0) { out[0] = out[1] = out[2] = 0.0; return; } /* :15-17 */
    SFX(hitrec) rec;
    if (SFX(hit_list)(w, o, d, (RT)1e-4, (RT)INFINITY, &rec)) { /* :19 */
        SFX(v3) att;
        if (g->mode == RTWO_RNG_PHILOX) rtwo_rng_next_event(g);  /* scatter at the e-th hit = event e */
        SFX(v3) nd = SFX(scatter)(w, g, d, &rec, &att);         /* :29 */
        double inner[3];
        SFX(ray_color)(w, g, rec.p, nd, depth - 1, inner);      /* :31 (s.reflected is always true, structs.jl:43) */
        out[0] = (double)att.x * inner[0];
        out[1] = (double)att.y * inner[1];
        out[2] = (double)att.z * inner[2];
    } else {
        SFX(skycolor)(d, out);                                  /* :36 */
    }
}

/* ---------------------------------------------------------------- camera, src/camera.jl:43-48 */

typedef struct {
    SFX(v3) origin, llc, horizontal, vertical, u, v, w;
    RT lens_radius;
} SFX(cam);

static inline void SFX(get_ray)(const SFX(cam)* c, rtwo_rng* g, RT s, RT t, SFX(v3)* o, SFX(v3)* d) {
    RT dx, dy;
    SFX(random_vec2_in_disk)(g, &dx, &dy); /* :44 -- always drawn, even for lens_radius == 0 */
    RT rx = c->lens_radius * dx, ry = c->lens_radius * dy;
    /* :45 offset = c.u*rd.x + c.v*rd.y */
    SFX(v3) off = SFX(mk)(FMA(c->v.x, ry, c->u.x * rx), FMA(c->v.y, ry, c->u.y * rx), FMA(c->v.z, ry, c->u.z * rx));
    *o = SFX(mk)(c->origin.x + off.x, c->origin.y + off.y, c->origin.z + off.z);
    /* :46-47 ((llc + s*horizontal) + t*vertical) - origin - offset */
    SFX(v3) q;
    q.x = FMA(t, c->vertical.x, FMA(s, c->horizontal.x, c->llc.x)) - c->origin.x - off.x;
    q.y = FMA(t, c->vertical.y, FMA(s, c->horizontal.y, c->llc.y)) - c->origin.y - off.y;
    q.z = FMA(t, c->vertical.z, FMA(s, c->horizontal.z, c->llc.z)) - c->origin.z - off.z;
    *d = SFX(normalize)(q);
}

/* ---------------------------------------------------------------- driver, src/render.jl:8-44 */

/* one (pixel, sample) of the loop body src/render.jl:26-38; i1,j1,s1 are the reference's 1-based indices */
static inline void SFX(sample_path)(SFX(world)* w, rtwo_rng* g, const SFX(cam)* c, int W, int H, int max_depth,
                                    int i1, int j1, int s1, double rgb[3]) {
    RT u = (RT)((double)j1 / (double)W);       /* :26 */
    RT v = (RT)((double)(H - i1) / (double)H); /* :27 */
    RT du = (RT)0, dv = (RT)0;                 /* :31 */
    if (s1 != 1) {
        rtwo_rng_seek(g, 0u, RT_WORDS_PER_DRAW); /* production stream: draws 0,1 of event 0 */
        du = SFX(trand)(g) / (RT)(float)W;     /* :34 (f32_image_width) */
        dv = SFX(trand)(g) / (RT)(float)H;     /* :35 */
    }
    SFX(v3) o, d;
    SFX(get_ray)(c, g, u + du, v + dv, &o, &d); /* :37 */
    SFX(ray_color)(w, g, o, d, max_depth, rgb); /* :38 */
}

typedef struct {
    const RT* geom4;
    const RT* mat4;
    const uint32_t* kind;
    uint32_t n;
    SFX(cam) cam;
    int W, H, spp, max_depth;
    uint64_t seed;
    int rng_mode;
    int row_start, row_stride;
    RT* out_rgb;
    double* out_linear;
    /* per worker */
    int tid, nthreads, joinable;
    uint64_t segments, tests;
} SFX(job);

static void SFX(render_pixel)(SFX(job)* jb, SFX(world)* w, rtwo_rng* g, int i0, int j0) {
    const int W = jb->W, H = jb->H;
    double acc[3] = {0.0, 0.0, 0.0}; /* :25 (promoted to Float64 by the first +=) */
    for (int s0 = 0; s0 < jb->spp; ++s0) { /* :29 */
        double rgb[3];
        if (jb->rng_mode == RTWO_RNG_PHILOX) rtwo_rng_begin_path(g, jb->seed, (uint32_t)(i0 * W + j0), (uint32_t)s0);
        SFX(sample_path)(w, g, &jb->cam, W, H, jb->max_depth, i0 + 1, j0 + 1, s0 + 1, rgb);
        acc[0] += rgb[0]; acc[1] += rgb[1]; acc[2] += rgb[2]; /* :38 */
    }
    size_t at = ((size_t)j0 * (size_t)H + (size_t)i0) * 3; /* Julia column-major img[i,j] */
    for (int k = 0; k < 3; ++k) {
        double lin = acc[k] / (double)jb->spp;      /* :40 accum_color / n_samples */
        if (jb->out_linear) jb->out_linear[at + k] = lin;
        jb->out_rgb[at + k] = (RT)sqrt(lin);        /* rgb_gamma2, src/vec.jl:22; store rounds to T */
    }
}

static void SFX(render_row)(SFX(job)* jb, SFX(world)* w, rtwo_rng* g, int i0) {
    for (int j0 = 0; j0 < jb->W; ++j0) SFX(render_pixel)(jb, w, g, i0, j0); /* :24 */
}

static void* SFX(worker)(void* arg) {
    SFX(job)* jb = (SFX(job)*)arg;
    SFX(world) w = {jb->geom4, jb->mat4, jb->kind, jb->n, 0, 0, NULL, 0, 0};
    rtwo_rng g;
    /* rows this call renders: r_k = row_start + k*row_stride, k = 0..nrows-1 */
    int nrows = (jb->H - jb->row_start + jb->row_stride - 1) / jb->row_stride;
    if (nrows < 0) nrows = 0;
    if (jb->rng_mode == RTWO_RNG_XOROSHIRO) {
        /* Threads.@threads static schedule (src/render.jl:23): contiguous row blocks, thread k uses TRNG[k]
         * reseeded with k (src/rand.jl:2) */
        rtwo_rng_seed_xoroshiro(&g, (uint64_t)(jb->tid + 1));
        int len = nrows / jb->nthreads, rem = nrows % jb->nthreads;
        int lo = jb->tid * len + (jb->tid < rem ? jb->tid : rem);
        int hi = lo + len + (jb->tid < rem ? 1 : 0);
        for (int k = lo; k < hi; ++k) SFX(render_row)(jb, &w, &g, jb->row_start + k * jb->row_stride);
    } else {
        memset(&g, 0, sizeof g);
        g.mode = RTWO_RNG_PHILOX;
        /* path-keyed stream: any pixel->thread map gives the same image; blocks of 16 pixels are dealt to the
         * threads round-robin (load balance, and single-row slices of a wide image still use every core) */
        const long total = (long)nrows * (long)jb->W;
        for (long b = jb->tid; b * 16 < total; b += jb->nthreads) {
            const long end = (b + 1) * 16 < total ? (b + 1) * 16 : total;
            for (long p = b * 16; p < end; ++p)
                SFX(render_pixel)(jb, &w, &g, jb->row_start + (int)(p / jb->W) * jb->row_stride, (int)(p % jb->W));
        }
    }
    jb->segments = w.segments;
    jb->tests = w.tests;
    return NULL;
}
