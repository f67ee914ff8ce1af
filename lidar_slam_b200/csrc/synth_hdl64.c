/*
 * synth_hdl64.c -- deterministic synthetic HDL-64-shaped workload generator (host C, no GPU).
 *
 * There is no network and the reference ships no point-cloud data (the map PCD is a missing large
 * blob), so tests and bench.py ray-cast a procedural street scene with the
 * sensor geometry SURVEY.md section 8(d) fixes: 64 rings from +2.0 deg to -24.8 deg, 2083 azimuth steps
 * per revolution (133 312 rays), sensor 1.73 m above the ground, 120 m max range, N(0, 0.02 m) range
 * noise, 10 % drop-outs => ~120 k returns per scan, intensity U[0,1].
 *
 * Everything is a pure function of (scene seed, frame id, ray id): counter-based splitmix64 hashing,
 * so scans can be generated in any order / on any number of threads and are bit-reproducible.
 * This is workload generation only; it is not part of the registration path and not the oracle.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_RINGS 64
#define N_AZ 2083
#define MAX_RANGE 120.0
#define SENSOR_Z 1.73

typedef struct { double lo[3], hi[3]; } box_t;
typedef struct { double cx, cy, r, h; } pole_t;

typedef struct {
    uint64_t seed;
    int n_box, n_pole;
    box_t *box;
    pole_t *pole;
    /* path: polyline with rounded corners, parametrised by arclength */
    double leg;      /* length of each straight leg (m) */
    double turn_r;   /* corner radius */
} scene_t;

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline uint64_t hash3(uint64_t a, uint64_t b, uint64_t c) {
    return splitmix64(splitmix64(splitmix64(a) ^ b) ^ c);
}
static inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
static inline double gauss(uint64_t h1, uint64_t h2) {
    double u1 = u01(h1), u2 = u01(h2);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

/* --- path: an L-shaped drive along two streets: leg along +x on the street y = 25, a quarter turn of radius
 * turn_r inside the intersection with the street x = 25 + leg, then a leg along +y on that street;
 * s in [0, 2*leg + arc] (the straight part of the first leg is leg - turn_r long) --- */
void synth_path_pose(const scene_t *sc, double s, double pose6[6]) {
    double L = sc->leg - sc->turn_r, r = sc->turn_r;
    double arc = 1.5707963267948966 * r;
    double x, y, yaw;
    if (s < 0) s = 0;
    if (s <= L) { x = 25.0 + s; y = 25.0; yaw = 0.0; }
    else if (s <= L + arc) {
        double a = (s - L) / r;
        x = 25.0 + L + r * sin(a);
        y = 25.0 + r * (1.0 - cos(a));
        yaw = a;
    } else {
        double t = s - L - arc;
        x = 25.0 + L + r; y = 25.0 + r + t; yaw = 1.5707963267948966;
    }
    /* lane weave + small roll / pitch so that all six degrees of freedom move */
    double w = 0.6 * sin(s * 0.05);
    x += -sin(yaw) * w;
    y += cos(yaw) * w;
    pose6[0] = x; pose6[1] = y; pose6[2] = SENSOR_Z;
    pose6[3] = 0.006 * sin(s * 0.11);        /* roll  ~ +-0.35 deg */
    pose6[4] = 0.008 * sin(s * 0.07 + 1.0);  /* pitch ~ +-0.45 deg */
    pose6[5] = yaw + 0.02 * sin(s * 0.03);
}
double synth_path_length(const scene_t *sc) { return 2.0 * sc->leg + 1.5707963267948966 * sc->turn_r; }

/* rotation R = Rx(roll) Ry(pitch) Rz(yaw) (the convention of the reference's 6-vector,
 * NormalDistributionsTransform.cpp:370-373), row-major double */
void synth_pose_to_matrix(const double p[6], double T[16]) {
    double cx = cos(p[3]), sx = sin(p[3]), cy = cos(p[4]), sy = sin(p[4]), cz = cos(p[5]), sz = sin(p[5]);
    double R[9] = {cy * cz, -cy * sz, sy,
                   cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy,
                   sx * sz - cx * sy * cz, sx * cz + cx * sy * sz, cx * cy};
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c]; T[r * 4 + 3] = p[r]; }
    T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
}

scene_t *synth_scene_create(uint64_t seed, double leg) {
    scene_t *sc = (scene_t *)calloc(1, sizeof(scene_t));
    sc->seed = seed;
    sc->leg = leg;
    sc->turn_r = 12.0;
    double span = leg + 25.0 + sc->turn_r;
    int lat_lo = -3, lat_hi = (int)ceil((span + 150.0) / 50.0);
    int nlat = lat_hi - lat_lo + 1;
    int max_box = nlat * nlat + 4096, max_pole = 4096;
    sc->box = (box_t *)malloc(max_box * sizeof(box_t));
    sc->pole = (pole_t *)malloc(max_pole * sizeof(pole_t));
    int nb = 0, np = 0;
    /* buildings on a 50 m lattice, centres jittered, 10-40 m wide, 5-20 m tall; the streets are the
     * lines y = 25 + 50k and x = 25 + 50k, which stay clear by at least 5 m */
    for (int i = lat_lo; i <= lat_hi; ++i)
        for (int j = lat_lo; j <= lat_hi; ++j) {
            uint64_t h = hash3(seed, (uint64_t)(i + 1000), (uint64_t)(j + 1000));
            double wx = 10.0 + 30.0 * u01(splitmix64(h ^ 1)), wy = 10.0 + 30.0 * u01(splitmix64(h ^ 2));
            double ht = 5.0 + 15.0 * u01(splitmix64(h ^ 3));
            double jx = (40.0 - wx) * 0.5 * (2.0 * u01(splitmix64(h ^ 4)) - 1.0);
            double jy = (40.0 - wy) * 0.5 * (2.0 * u01(splitmix64(h ^ 5)) - 1.0);
            double cx = 50.0 * i + jx, cy = 50.0 * j + jy;
            box_t *b = &sc->box[nb++];
            b->lo[0] = cx - wx / 2; b->hi[0] = cx + wx / 2;
            b->lo[1] = cy - wy / 2; b->hi[1] = cy + wy / 2;
            b->lo[2] = 0.0; b->hi[2] = ht;
        }
    /* parked cars (4.2 x 1.8 x 1.5 m) and poles along the driven corridor */
    double plen = synth_path_length(sc);
    /* the quarter turn occupies s in [s_turn0 - margin, s_turn1 + margin]: keep kerb-side objects out of it (they
     * are offset sideways from the centre line with the heading of the straight they stand on) */
    double s_turn0 = sc->leg - sc->turn_r - 6.0, s_turn1 = sc->leg - sc->turn_r + 1.5707963267948966 * sc->turn_r + 6.0;
    int ncar = (int)(plen / 14.0), npol = (int)(plen / 18.0);
    for (int k = 0; k < ncar && nb < max_box; ++k) {
        uint64_t h = hash3(seed, 0xCA5, (uint64_t)k);
        double s = plen * (k + u01(splitmix64(h ^ 1))) / ncar;
        double p[6];
        /* un-weaved centre line */
        scene_t tmp = *sc;
        synth_path_pose(&tmp, s, p);
        double yaw = (s <= s_turn0) ? 0.0 : ((s >= s_turn1) ? 1.5707963267948966 : -1.0);
        if (yaw < 0) continue; /* no cars in the corner */
        double side = (splitmix64(h ^ 2) & 1) ? 1.0 : -1.0;
        double off = side * (3.2 + 0.8 * u01(splitmix64(h ^ 3)));
        double cx = p[0] - sin(yaw) * off, cy = p[1] + cos(yaw) * off;
        double lx = (yaw == 0.0) ? 4.2 : 1.8, ly = (yaw == 0.0) ? 1.8 : 4.2;
        box_t *b = &sc->box[nb++];
        b->lo[0] = cx - lx / 2; b->hi[0] = cx + lx / 2;
        b->lo[1] = cy - ly / 2; b->hi[1] = cy + ly / 2;
        b->lo[2] = 0.0; b->hi[2] = 1.4 + 0.3 * u01(splitmix64(h ^ 4));
    }
    for (int k = 0; k < npol && np < max_pole; ++k) {
        uint64_t h = hash3(seed, 0x901E, (uint64_t)k);
        double s = plen * (k + u01(splitmix64(h ^ 1))) / npol;
        double p[6];
        synth_path_pose(sc, s, p);
        if (s > s_turn0 && s < s_turn1) continue; /* no poles in the corner */
        double yaw = (s <= s_turn0) ? 0.0 : 1.5707963267948966;
        double side = (splitmix64(h ^ 2) & 1) ? 1.0 : -1.0;
        double off = side * (4.4 + 0.4 * u01(splitmix64(h ^ 3)));
        pole_t *q = &sc->pole[np++];
        q->cx = p[0] - sin(yaw) * off; q->cy = p[1] + cos(yaw) * off;
        q->r = 0.15; q->h = 5.0 + 4.0 * u01(splitmix64(h ^ 4));
    }
    sc->n_box = nb; sc->n_pole = np;
    return sc;
}
/* 1 when (x, y) lies inside the footprint of any box of the scene (building or car) grown by `margin` */
int synth_point_in_box(const scene_t *sc, double x, double y, double margin) {
    for (int i = 0; i < sc->n_box; ++i)
        if (x > sc->box[i].lo[0] - margin && x < sc->box[i].hi[0] + margin && y > sc->box[i].lo[1] - margin && y < sc->box[i].hi[1] + margin)
            return 1;
    return 0;
}
void synth_scene_free(scene_t *sc) { if (sc) { free(sc->box); free(sc->pole); free(sc); } }

/* nearest hit along origin o + t d, t in (0, MAX_RANGE]; returns t or -1 */
static double cast_ray(const scene_t *sc, const int *bidx, int nbi, const int *pidx, int npi,
                       const double o[3], const double d[3]) {
    double best = MAX_RANGE + 1.0;
    if (d[2] < 0.0) { double t = -o[2] / d[2]; if (t > 0 && t < best) best = t; }
    for (int k = 0; k < nbi; ++k) {
        const box_t *b = &sc->box[bidx[k]];
        double t0 = 0.0, t1 = best;
        int ok = 1;
        for (int a = 0; a < 3 && ok; ++a) {
            if (fabs(d[a]) < 1e-12) { if (o[a] < b->lo[a] || o[a] > b->hi[a]) ok = 0; }
            else {
                double inv = 1.0 / d[a];
                double ta = (b->lo[a] - o[a]) * inv, tb = (b->hi[a] - o[a]) * inv;
                if (ta > tb) { double t = ta; ta = tb; tb = t; }
                if (ta > t0) t0 = ta;
                if (tb < t1) t1 = tb;
                if (t0 > t1) ok = 0;
            }
        }
        if (ok && t0 > 1e-6 && t0 < best) best = t0;
    }
    for (int k = 0; k < npi; ++k) {
        const pole_t *q = &sc->pole[pidx[k]];
        double ox = o[0] - q->cx, oy = o[1] - q->cy;
        double A = d[0] * d[0] + d[1] * d[1];
        if (A < 1e-14) continue;
        double B = ox * d[0] + oy * d[1], Cc = ox * ox + oy * oy - q->r * q->r;
        double disc = B * B - A * Cc;
        if (disc < 0) continue;
        double t = (-B - sqrt(disc)) / A;
        if (t > 1e-6 && t < best) {
            double z = o[2] + t * d[2];
            if (z >= 0.0 && z <= q->h) best = t;
        }
    }
    return best <= MAX_RANGE ? best : -1.0;
}

static void cull(const scene_t *sc, const double o[3], int *bidx, int *nbi, int *pidx, int *npi) {
    int nb = 0, np = 0;
    for (int k = 0; k < sc->n_box; ++k) {
        const box_t *b = &sc->box[k];
        double dx = fmax(fmax(b->lo[0] - o[0], 0.0), o[0] - b->hi[0]);
        double dy = fmax(fmax(b->lo[1] - o[1], 0.0), o[1] - b->hi[1]);
        if (dx * dx + dy * dy <= MAX_RANGE * MAX_RANGE) bidx[nb++] = k;
    }
    for (int k = 0; k < sc->n_pole; ++k) {
        double dx = sc->pole[k].cx - o[0], dy = sc->pole[k].cy - o[1];
        if (dx * dx + dy * dy <= (MAX_RANGE + 1) * (MAX_RANGE + 1)) pidx[np++] = k;
    }
    *nbi = nb; *npi = np;
}

/* One scan at sensor pose6 (x,y,z,roll,pitch,yaw in the map frame).  keep_prob < 1 thins the rays
 * (used when assembling maps).  world != 0 writes points in the map frame, else in the sensor frame.
 * out: capacity N_RINGS*N_AZ*4 floats (x,y,z,intensity).  Returns number of returns. */
size_t synth_scan(const scene_t *sc, uint64_t frame_id, const double pose6[6], double keep_prob, int world,
                  float *out) {
    double T[16];
    synth_pose_to_matrix(pose6, T);
    double o[3] = {T[3], T[7], T[11]};
    int *bidx = (int *)malloc(sizeof(int) * (sc->n_box + 1));
    int *pidx = (int *)malloc(sizeof(int) * (sc->n_pole + 1));
    int nbi, npi;
    cull(sc, o, bidx, &nbi, pidx, &npi);
    uint64_t fseed = hash3(sc->seed, 0x5EED0000ULL + frame_id, 0x11);
    size_t n = 0;
    for (int a = 0; a < N_AZ; ++a) {
        double az = a * (6.283185307179586 / N_AZ);
        double ca = cos(az), sa = sin(az);
        for (int r = 0; r < N_RINGS; ++r) {
            uint64_t rid = (uint64_t)a * N_RINGS + r;
            uint64_t h = splitmix64(fseed ^ (rid * 0x9E3779B97F4A7C15ULL));
            if (keep_prob < 1.0 && u01(splitmix64(h ^ 0xA1)) >= keep_prob) continue;
            if (u01(splitmix64(h ^ 0xB2)) < 0.10) continue; /* drop-out */
            double el = (2.0 - r * (26.8 / 63.0)) * (3.141592653589793 / 180.0);
            double ce = cos(el), se = sin(el);
            double ds[3] = {ce * ca, ce * sa, se};
            double d[3];
            for (int k = 0; k < 3; ++k) d[k] = T[k * 4 + 0] * ds[0] + T[k * 4 + 1] * ds[1] + T[k * 4 + 2] * ds[2];
            double t = cast_ray(sc, bidx, nbi, pidx, npi, o, d);
            if (t < 0) continue;
            t += 0.02 * gauss(splitmix64(h ^ 0xC3), splitmix64(h ^ 0xD4));
            if (t < 0.5) continue;
            float *p = &out[4 * n];
            if (world) { for (int k = 0; k < 3; ++k) p[k] = (float)(o[k] + t * d[k]); }
            else       { for (int k = 0; k < 3; ++k) p[k] = (float)(t * ds[k]); }
            p[3] = (float)u01(splitmix64(h ^ 0xE5));
            ++n;
        }
    }
    free(bidx); free(pidx);
    return n;
}

/* many scans on worker threads; out_offsets[i] = i * (N_RINGS*N_AZ) points; counts[i] returned */
typedef struct {
    const scene_t *sc; const uint64_t *ids; const double *poses; int n; int stride_pts;
    float *out; size_t *counts; int tid, nthreads;
} job_t;
static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (int i = j->tid; i < j->n; i += j->nthreads)
        j->counts[i] = synth_scan(j->sc, j->ids[i], &j->poses[6 * i], 1.0, 0, j->out + (size_t)i * j->stride_pts * 4);
    return NULL;
}
void synth_scans(const scene_t *sc, const uint64_t *ids, const double *poses6, int n, int stride_pts,
                 float *out, size_t *counts, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    job_t jb[256];
    for (int t = 0; t < nthreads; ++t) {
        jb[t] = (job_t){sc, ids, poses6, n, stride_pts, out, counts, t, nthreads};
        pthread_create(&th[t], NULL, worker, &jb[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* Map: union of thinned scans taken every `spacing` metres along the path, in the map frame,
 * then reduced to exactly n_points by keeping the n_points smallest point hashes (a uniform random
 * subset that does not depend on generation order).  Returns the number of points written. */
typedef struct { uint64_t h; uint32_t i; } hk_t;
static int cmp_hk(const void *a, const void *b) {
    const hk_t *x = (const hk_t *)a, *y = (const hk_t *)b;
    return x->h < y->h ? -1 : (x->h > y->h);
}
size_t synth_map(const scene_t *sc, size_t n_points, double spacing, float *out, int nthreads) {
    int nscan = (int)floor(synth_path_length(sc) / spacing) + 1;
    double per_scan = 1.15 * (double)n_points / nscan;
    double keep = per_scan / (0.9 * 0.93 * N_RINGS * N_AZ);
    if (keep > 1.0) keep = 1.0;
    /* generate in chunks to bound memory: each scan writes at most keep*rays (+ slack) points */
    size_t cap_per = (size_t)(keep * N_RINGS * N_AZ * 1.3) + 4096;
    if (cap_per > (size_t)N_RINGS * N_AZ) cap_per = (size_t)N_RINGS * N_AZ;
    float *all = (float *)malloc((size_t)nscan * cap_per * 4 * sizeof(float));
    (void)nthreads;
    float *tmp = (float *)malloc((size_t)N_RINGS * N_AZ * 4 * sizeof(float));
    size_t total = 0;
    for (int i = 0; i < nscan; ++i) {
        double p[6];
        synth_path_pose(sc, i * spacing, p);
        size_t c = synth_scan(sc, 0x4D415000ULL + (uint64_t)i, p, keep, 1, tmp);
        if (c > cap_per) c = cap_per;
        memcpy(all + total * 4, tmp, c * 4 * sizeof(float));
        total += c;
    }
    free(tmp);
    size_t keepn = total < n_points ? total : n_points;
    if (total > n_points) {
        hk_t *hk = (hk_t *)malloc(total * sizeof(hk_t));
        for (size_t i = 0; i < total; ++i) { hk[i].h = hash3(sc->seed, 0x5E1EC7, i); hk[i].i = (uint32_t)i; }
        qsort(hk, total, sizeof(hk_t), cmp_hk);
        uint64_t thr = hk[n_points - 1].h;
        free(hk);
        size_t w = 0;
        for (size_t i = 0; i < total && w < n_points; ++i)
            if (hash3(sc->seed, 0x5E1EC7, i) <= thr) { memcpy(out + 4 * w, all + 4 * i, 4 * sizeof(float)); ++w; }
        keepn = w;
    } else {
        memcpy(out, all, total * 4 * sizeof(float));
    }
    free(all);
    return keepn;
}

int synth_rays_per_scan(void) { return N_RINGS * N_AZ; }
