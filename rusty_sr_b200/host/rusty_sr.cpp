// rusty_sr.cpp — compiled host program mirroring the reference CLI (src/main.rs:33-178) on top of the
// C ABI (include/rsr_b200.h).  The reference's host is Rust; there is no Rust toolchain in this image, so
// the mirror is C++: same positional arguments and flags (clap definitions main.rs:39-71), same progress
// messages (main.rs:137-177), same parameter handling (.rsr decode = bytevec <Vec<f32>>::decode::<u32>,
// main.rs:138-152; size check main.rs:162), with `graph.forward` (main.rs:171) replaced by rsr_upscale_u8.
//
//   rusty_sr INPUT_FILE OUTPUT_FILE [-p imagenet|imagenetlinear|anime|bilinear] [-c PARAMETER_FILE] [-d]
//
// PNG only (host I/O; the reference reads whatever the `image` crate reads).  `-p bilinear` and `-d` do not
// use the conv net (network.rs:111-138); they run as small elementwise kernels (rsr_bilinear_u8 / rsr_downsample_u8).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rsr_b200.h"
#include "png_io.hpp"

static int FACTOR = 3;  // main.rs:31 (`const FACTOR: usize = 3; // TODO expose as an option`): --factor exposes it

// bytevec decode::<u32> of Vec<f32> (main.rs:138,146-152) through the library's codec (rsr_rsr_decode)
static std::vector<float> decode_rsr(const std::vector<uint8_t>& d) {
    const long long n = rsr_rsr_decode(d.data(), d.size(), nullptr, 0);
    if (n < 0) throw std::runtime_error("ByteVec conversion failed");  // main.rs:138 .expect(..)
    std::vector<float> out((size_t)n);
    rsr_rsr_decode(d.data(), d.size(), out.data(), out.size());
    return out;
}
// bytevec encode::<u32> (main.rs:213: the checkpoint writer of the training loop)
static void write_rsr(const std::string& path, const std::vector<float>& params) {
    std::vector<uint8_t> blob(rsr_rsr_encoded_size(params.size()));
    if (rsr_rsr_encode(params.data(), params.size(), blob.data(), blob.size()) < 0) throw std::runtime_error(rsr_last_error());
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(blob.data(), 1, blob.size(), f) != blob.size()) { if (f) fclose(f); throw std::runtime_error("Error writing parameter file"); }
    fclose(f);
}

static std::vector<uint8_t> read_file(const std::string& path, const char* what) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error(what);
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return buf;
}

static std::string resource_dir(const char* argv0) {
    if (const char* e = getenv("RUSTY_SR_RES")) return e;
    std::string p = argv0;
    const size_t s = p.find_last_of('/');
    p = s == std::string::npos ? "." : p.substr(0, s);
    return p + "/../res";  // rusty_sr_b200/bin/rusty_sr -> rusty_sr_b200/res (the blobs main.rs:26-28 embeds)
}

static int usage(const char* msg) {
    fprintf(stderr, "error: %s\n\nUSAGE:\n    rusty_sr <INPUT_FILE> <OUTPUT_FILE> [-p <PARAMETERS>] [-c <PARAMETER_FILE>] [-d]\n"
                    "    PARAMETERS: imagenet (default), imagenetlinear, anime, bilinear\n", msg);
    return 2;
}

int main(int argc, char** argv) {
    std::vector<std::string> pos;
    std::string parameters, custom, export_params;
    bool downsample = false;
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](const char* name) -> std::string {
            if (i + 1 >= argc) { usage((std::string("missing value for ") + name).c_str()); exit(2); }
            return argv[++i];
        };
        if (a == "-p" || a == "--parameters") parameters = value("--parameters");
        else if (a == "-c" || a == "--custom") custom = value("--custom");
        else if (a == "-d" || a == "--downsample") downsample = true;
        else if (a == "--device") device = atoi(value("--device").c_str());
        else if (a == "--factor") FACTOR = atoi(value("--factor").c_str());             // extension (reference: compile-time 3)
        else if (a == "--export-params") export_params = value("--export-params");       // extension: re-encode the selected .rsr
        else if (a == "--png-roundtrip") {  // test hook for the PNG codec: decode INPUT, re-encode as RGBA
            if (i + 2 >= argc) return usage("--png-roundtrip IN OUT");
            try {
                const rsr_host::Image im = rsr_host::read_png(argv[i + 1]);
                rsr_host::write_png_rgba(argv[i + 2], im.rgba.data(), im.w, im.h);
            } catch (const std::exception& e) { fprintf(stderr, "rusty_sr: %s\n", e.what()); return 101; }
            return 0;
        }
        else if (!a.empty() && a[0] == '-') return usage(("unknown flag " + a).c_str());
        else pos.push_back(a);
    }
    if (pos.size() != 2) return usage("INPUT_FILE and OUTPUT_FILE are required");
    if (FACTOR < 1 || FACTOR > 4) return usage("--factor must be 1..4");
    // clap conflicts_with (main.rs:58-70)
    if (!custom.empty() && !parameters.empty()) return usage("--custom conflicts with --parameters");
    if (downsample && (!custom.empty() || !parameters.empty())) return usage("--downsample conflicts with --parameters / --custom");
    if (!parameters.empty() && parameters != "imagenet" && parameters != "imagenetlinear" && parameters != "anime" &&
        parameters != "bilinear")
        return usage("possible values for --parameters: imagenet, imagenetlinear, anime, bilinear");
    try {
        std::vector<float> params;
        if (!custom.empty()) {
            const auto data = read_file(custom, "Error opening parameter file");  // main.rs:134
            printf("Upscaling using custom neural net parameters...");
            params = decode_rsr(data);
        } else if (downsample || parameters == "bilinear") {
            // parameter-free graphs: downsample_net / bilinear_net (network.rs:111-138, main.rs:139-141,153-155)
            printf(downsample ? "Downsampling using average pooling of linear RGB values..." : "Upscaling using bilinear interpolation...");
            fflush(stdout);
            const rsr_host::Image in = rsr_host::read_png(pos[0]);
            const int ow = downsample ? (in.w + FACTOR - 1) / FACTOR : in.w * FACTOR;
            const int oh = downsample ? (in.h + FACTOR - 1) / FACTOR : in.h * FACTOR;
            std::vector<uint8_t> out((size_t)ow * oh * 4);
            const int rc = downsample ? rsr_downsample_u8(device, in.rgba.data(), in.w, in.h, 4, FACTOR, out.data())
                                      : rsr_bilinear_u8(device, in.rgba.data(), in.w, in.h, 4, FACTOR, out.data());
            if (rc) throw std::runtime_error(rsr_last_error());
            printf(" Writing file...");
            fflush(stdout);
            rsr_host::write_png_rgba(pos[1], out.data(), ow, oh);
            printf(" Done\n");
            return 0;
        } else {
            const std::string name = parameters.empty() ? "imagenet" : parameters;  // default main.rs:144
            printf("Upscaling using %sneural net parameters...", name == "imagenet" ? "imagenet " : name == "anime" ? "anime "
                                                                 : "linear loss imagenet ");
            params = decode_rsr(read_file(resource_dir(argv[0]) + "/" + name + ".rsr", "built-in parameter blob not found (set RUSTY_SR_RES)"));
        }
        fflush(stdout);
        if (!export_params.empty()) write_rsr(export_params, params);
        // main.rs:162 assert_eq!(params.len(), graph.num_params(), ...)
        if (params.size() != rsr_num_params(FACTOR))
            throw std::runtime_error("Parameters selected do not have the size required by the neural net. Ensure that the "
                                     "same sample factor is used for upscaling and training");
        const rsr_host::Image in = rsr_host::read_png(pos[0]);
        rsr_ctx* ctx = rsr_create(params.data(), params.size(), FACTOR, device, 0);
        if (!ctx) throw std::runtime_error(rsr_last_error());
        // the output image lives in page-locked memory: the result's copy back then runs at PCIe speed and overlaps the computation
        const size_t out_bytes = (size_t)in.w * FACTOR * in.h * FACTOR * 4;
        uint8_t* out = static_cast<uint8_t*>(rsr_host_alloc(out_bytes));
        if (!out) { const std::string e = rsr_last_error(); rsr_destroy(ctx); throw std::runtime_error(e); }
        // img_to_data -> graph.forward -> data_to_img(..).to_rgba()   (main.rs:170-175), fused on the device
        const int rc = rsr_upscale_u8(ctx, in.rgba.data(), in.w, in.h, 4, out);
        const std::string err = rc ? rsr_last_error() : "";
        rsr_destroy(ctx);
        if (rc) { rsr_host_free(out); throw std::runtime_error(err); }
        printf(" Writing file...");
        fflush(stdout);
        try {
            rsr_host::write_png_rgba(pos[1], out, in.w * FACTOR, in.h * FACTOR);
        } catch (...) {
            rsr_host_free(out);
            throw;
        }
        rsr_host_free(out);
        printf(" Done\n");
    } catch (const std::exception& e) {
        fprintf(stderr, "\nrusty_sr: %s\n", e.what());  // the reference panics here
        return 101;                                    // Rust's panic exit code
    }
    return 0;
}
