// nff.cpp -- the reference's NFF (Neutral File Format, SPD) reader, GlomeTrace/Data/Glome/Spd.hs:1-261,
// restated over the host scene-construction mirror.
//
// What Spd.hs does, quirks included, because the scene a file yields is part of the reference's behaviour:
//   - the file is a sequence of cameras ("v" ... "resolution" x y), fill groups, lights ("l") and backgrounds
//     ("b"), tried in that order at each position (accum_rss, Spd.hs:199-234); parsing stops silently at the first
//     position where none of them parses;
//   - a fill group is an "f" line followed by as many primitives as parse ("s", "c", "p", "pp"); it becomes
//     tex (bih prims) (Surface clr (1-T) 0 kd ks shine)  (readsSpdFill / readsSpdTextureGroup, Spd.hs:137-196);
//     primitives that do not follow an "f" end the parse;
//   - "p n" / "pp n": n is read and ignored, vertices are read greedily, the polygon is a triangle fan
//     (Triangle.hs:29-44) wrapped in `group`;
//   - every list is accumulated by consing, so groups and lights end up in REVERSE file order, and the scene uses
//     the LAST camera and the LAST background; a file with no camera or no background is a pattern-match failure
//     (readsSpdScene, Spd.hs:243-246) -> BuildError here;
//   - `lexcr` treats a "#" token by discarding tokens to the end of the INPUT, not of the line (Haskell's `lex`
//     skips newlines as white space, Spd.hs:12-29), so a comment ends the scene.  Reproduced.
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_builder.h"

namespace glome_host {
namespace {

struct Lexer {
    const char* p;
    const char* end;
    // next white-space separated token without consuming it; false at end of input or at a comment
    bool peek(std::string& tok, const char** after) const {
        const char* q = p;
        while (q < end && isspace((unsigned char)*q)) q++;
        if (q >= end) return false;
        const char* s = q;
        while (q < end && !isspace((unsigned char)*q)) q++;
        tok.assign(s, q);
        if (tok[0] == '#') {
            // Haskell's lex reads a maximal run of symbol characters: "#" followed by a non-symbol is the comment token
            static const char* sym = "!@#$%&*+./<=>?\\^|:-~";
            if (tok.size() == 1 || !strchr(sym, tok[1])) return false;
        }
        *after = q;
        return true;
    }
    bool word(const char* w) {
        std::string t; const char* a;
        if (!peek(t, &a) || t != w) return false;
        p = a;
        return true;
    }
    bool skip() {  // any one token (hither, resolution values)
        std::string t; const char* a;
        if (!peek(t, &a)) return false;
        p = a;
        return true;
    }
    // reads :: Flt  -- digits[.digits][e[+-]digits] with an optional leading '-'
    bool number(Flt& out) {
        std::string t; const char* a;
        if (!peek(t, &a)) return false;
        size_t i = 0, n = t.size();
        if (i < n && t[i] == '-') i++;
        size_t d0 = i;
        while (i < n && isdigit((unsigned char)t[i])) i++;
        if (i == d0) return false;
        if (i < n && t[i] == '.') {
            i++;
            size_t f0 = i;
            while (i < n && isdigit((unsigned char)t[i])) i++;
            if (i == f0) return false;
        }
        if (i < n && (t[i] == 'e' || t[i] == 'E')) {
            i++;
            if (i < n && (t[i] == '-' || t[i] == '+')) i++;
            size_t e0 = i;
            while (i < n && isdigit((unsigned char)t[i])) i++;
            if (i == e0) return false;
        }
        if (i != n) return false;
        out = strtod(t.c_str(), nullptr);
        p = a;
        return true;
    }
    bool integer() {  // reads :: Int
        std::string t; const char* a;
        if (!peek(t, &a)) return false;
        size_t i = 0;
        if (i < t.size() && t[i] == '-') i++;
        if (i == t.size()) return false;
        for (; i < t.size(); i++) if (!isdigit((unsigned char)t[i])) return false;
        p = a;
        return true;
    }
    bool vec3(Vec& v) {
        Lexer save = *this;
        if (number(v.x) && number(v.y) && number(v.z)) return true;
        *this = save;
        return false;
    }
};

// one primitive of readsSpdSolid (Spd.hs:158-176); false = none parses here
bool parse_solid(Builder& b, Lexer& lx, int& item) {
    Lexer save = lx;
    if (lx.word("s")) {
        Vec c; Flt r;
        if (lx.vec3(c) && lx.number(r)) { item = b.sphere(c, r); return true; }
        lx = save;
        return false;
    }
    if (lx.word("c")) {
        Vec e1, e2; Flt r1, r2;
        if (lx.vec3(e1) && lx.number(r1) && lx.vec3(e2) && lx.number(r2)) { item = b.cone(e1, r1, e2, r2); return true; }
        lx = save;
        return false;
    }
    if (lx.word("p")) {
        if (!lx.integer()) { lx = save; return false; }
        std::vector<Vec> vs;
        Vec v;
        while (lx.vec3(v)) vs.push_back(v);  // readsSpdVecs: greedy, the count is ignored
        if (vs.empty()) throw BuildError("nff: polygon without vertices (triangles [] has no equation, Triangle.hs:30)");
        std::vector<int32_t> tris;
        for (size_t i = 1; i + 1 < vs.size(); i++) tris.push_back(b.triangle(vs[0], vs[i], vs[i + 1]));
        item = b.group(tris);
        return true;
    }
    if (lx.word("pp")) {
        if (!lx.integer()) { lx = save; return false; }
        std::vector<Vec> vs, ns;
        for (;;) {
            Lexer s2 = lx;
            Vec v, n;
            if (lx.vec3(v) && lx.vec3(n)) { vs.push_back(v); ns.push_back(n); }
            else { lx = s2; break; }
        }
        if (vs.empty()) throw BuildError("nff: polygonal patch without vertices (Triangle.hs:40)");
        std::vector<int32_t> tris;
        for (size_t i = 1; i + 1 < vs.size(); i++) tris.push_back(b.trianglenorm(vs[0], vs[i], vs[i + 1], ns[0], ns[i], ns[i + 1]));
        item = b.group(tris);
        return true;
    }
    return false;
}

}  // namespace

int nff_load(Builder& b, const char* text, int64_t len, GlomeCamera* cam_out, double bg_out[3], int64_t* consumed) {
    Lexer lx{text, text + (len < 0 ? (int64_t)strlen(text) : len)};
    std::vector<int32_t> groups;
    struct L { Vec pos; Flt r, g, bl; };
    std::vector<L> lights;
    bool have_cam = false, have_bg = false;
    GlomeCamera cam;
    memset(&cam, 0, sizeof(cam));
    Flt bg[3] = {0, 0, 0};
    for (;;) {
        Lexer save = lx;
        // camera (readsSpdCam, Spd.hs:91-106)
        if (lx.word("v")) {
            Vec from, at, up; Flt angle;
            if (lx.word("from") && lx.vec3(from) && lx.word("at") && lx.vec3(at) && lx.word("up") && lx.vec3(up) &&
                lx.word("angle") && lx.number(angle) && lx.word("hither") && lx.skip() && lx.word("resolution") && lx.skip() && lx.skip()) {
                make_camera(from, at, up, angle, &cam);
                have_cam = true;
                continue;
            }
            lx = save;
            break;
        }
        // fill group (readsSpdTextureGroup, Spd.hs:189-193)
        if (lx.word("f")) {
            Flt r, g, bl, kd, ks, shine, trans, ior;
            if (lx.number(r) && lx.number(g) && lx.number(bl) && lx.number(kd) && lx.number(ks) && lx.number(shine) &&
                lx.number(trans) && lx.number(ior)) {
                int mat = b.mat_surface(r, g, bl, 1 - trans, 0, kd, ks, shine);  // Surface clr (1-trans) 0 kd ks shine False
                int t = b.tex_uniform(mat);
                std::vector<int32_t> prims;
                int item;
                while (parse_solid(b, lx, item)) prims.push_back(item);
                groups.push_back(b.tex(b.bih(prims), t));
                continue;
            }
            lx = save;
            break;
        }
        // light (readsSpdLight, Spd.hs:124-130): with a colour if three more numbers follow
        if (lx.word("l")) {
            L l;
            if (lx.vec3(l.pos)) {
                Vec c;
                if (lx.vec3(c)) { l.r = c.x; l.g = c.y; l.bl = c.z; }
                else { l.r = l.g = l.bl = 1; }
                lights.push_back(l);
                continue;
            }
            lx = save;
            break;
        }
        // background (readsSpdBackground, Spd.hs:116-119)
        if (lx.word("b")) {
            Vec c;
            if (lx.vec3(c)) { bg[0] = c.x; bg[1] = c.y; bg[2] = c.z; have_bg = true; continue; }
            lx = save;
            break;
        }
        break;
    }
    if (consumed) *consumed = (int64_t)(lx.p - text);
    if (!have_cam) throw BuildError("nff: no camera (readsSpdScene's pattern fails, Spd.hs:245)");
    if (!have_bg) throw BuildError("nff: no background colour (readsSpdScene's pattern fails, Spd.hs:245)");
    // every list was built by consing: reverse file order (accum_rss, Spd.hs:213-230)
    std::vector<int32_t> rev(groups.rbegin(), groups.rend());
    for (size_t i = lights.size(); i-- > 0;) b.light(lights[i].pos, lights[i].r, lights[i].g, lights[i].bl);
    if (cam_out) *cam_out = cam;
    if (bg_out) { bg_out[0] = bg[0]; bg_out[1] = bg[1]; bg_out[2] = bg[2]; }
    return b.bih(rev);  // SPD (bih prims) lights cam bgc
}

}  // namespace glome_host
