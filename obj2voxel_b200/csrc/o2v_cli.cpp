// obj2voxel-b200 — command line front end with the reference CLI's arguments (src/main.cpp:264-380, option texts
// src/constants.hpp:26-61), implemented only on top of the public C API of include/obj2voxel.h:
//
//   obj2voxel-b200 INPUT_FILE OUTPUT_FILE -r RES [-i obj|stl] [-o ply|qef|vl32|vox|xyzrgb] [-t TEXTURE]
//                  [-s max|blend] [-p PERMUTATION] [-u] [-j THREADS] [-v] [-V] [-h] [--80]
//
// Behaviour kept from the reference: both positionals and -r are required (otherwise the help text is printed and the
// exit status is 1), -V prints version information, -p takes three letters out of xyz / XYZ (a capital flips the axis),
// -u doubles the sampling resolution, -j starts that many obj2voxel_run_worker threads (the GPU does the voxelization;
// the workers only keep the reference's threading contract alive).  Stated deviation: the exit status is the
// obj2voxel_error_t of the job (the reference's main() drops mainImpl's result and always exits 0, src/main.cpp:362-373).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "obj2voxel.h"

namespace {

const char *kVersion = "1.3.5-dev (obj2voxel_b200: sm_100a CUDA hot path)";

struct Options {
    std::string input, output, inputFormat, outputFormat, texture, permutation = "xyz";
    unsigned resolution = 0;
    unsigned threads = 0;
    bool haveResolution = false, supersample = false, verbose = false, version = false, help = false, eighty = false;
    obj2voxel_enum_t strategy = OBJ2VOXEL_MAX_STRATEGY;
    bool bad = false;
};

void printWrapped(const char *flags, const char *text, unsigned width)
{
    const unsigned gutter = 30;
    std::string line = std::string("  ") + flags;
    if (line.size() + 2 > gutter) {
        printf("%s\n", line.c_str());
        line.clear();
    }
    line.resize(gutter, ' ');
    std::string word;
    std::string rest = text;
    size_t pos = 0;
    while (pos <= rest.size()) {
        const size_t space = rest.find(' ', pos);
        word = rest.substr(pos, space == std::string::npos ? std::string::npos : space - pos);
        if (line.size() + word.size() + 1 > width && line.size() > gutter) {
            printf("%s\n", line.c_str());
            line.assign(gutter, ' ');
        }
        line += word;
        line += ' ';
        if (space == std::string::npos) {
            break;
        }
        pos = space + 1;
    }
    printf("%s\n", line.c_str());
}

void printHelp(bool eighty)
{
    const unsigned width = eighty ? 80 : 120;
    printf("Usage: obj2voxel-b200 {OPTIONS} [INPUT_FILE] [OUTPUT_FILE]\n\n");
    printf("General Options:\n");
    printWrapped("-h, --help", "Display this help menu.", width);
    printWrapped("--80", "Print help menu in 80 column mode.", width);
    printWrapped("-v, --verbose", "Enables verbose logging.", width);
    printWrapped("-V, --version", "Displays the version and other information.", width);
    printf("File Options:\n");
    printWrapped("INPUT_FILE", "First argument. Path to input file.", width);
    printWrapped("OUTPUT_FILE", "Second argument. Path to output file.", width);
    printWrapped("-i[obj|stl]", "Explicit input format. (Optional)", width);
    printWrapped("-o[ply|qef|vl32|vox|xyzrgb]", "Explicit output format. (Optional)", width);
    printWrapped("-t[texture]",
                 "Fallback texture path. Used when model has UV coordinates but textures can't be found in the material "
                 "library. (Default: none)",
                 width);
    printf("Voxelization Options:\n");
    printWrapped("-r[resolution], --res=[resolution]", "Maximum voxel grid resolution on any axis. (Required)", width);
    printWrapped("-s[max|blend], --strat=[max|blend]",
                 "Strategy for combining voxels of different triangles. Blend gives smoother colors at triangle edges "
                 "but might produce new and unwanted colors. (Default: max)",
                 width);
    printWrapped("-p[permutation], --perm=[permutation]",
                 "Permutation of xyz axes in the model. Capital letters flip an axis. (e.g. xYz to flip y-axis) "
                 "(Default: xyz)",
                 width);
    printWrapped("-u, --super",
                 "Enables supersampling. The model is voxelized at double resolution and then downscaled while "
                 "combining colors.",
                 width);
    printWrapped("-j[threads], --threads=[threads]",
                 "Number of worker threads to be started. The voxelization itself runs on the GPU; worker threads only "
                 "keep the library's threading contract. (Default: 0)",
                 width);
    printf("Visit at https://github.com/eisenwave/obj2voxel\n");
}

/// Value of a flag given as "-r 64", "-r64", "--res=64" or "--res 64".
bool takeValue(int argc, char **argv, int &i, const char *shortFlag, const char *longFlag, std::string *out)
{
    const std::string arg = argv[i];
    const std::string s = shortFlag, l = std::string("--") + longFlag;
    if (arg == s || (longFlag[0] != 0 && arg == l)) {
        if (i + 1 >= argc) {
            return false;
        }
        *out = argv[++i];
        return true;
    }
    if (arg.compare(0, s.size(), s) == 0 && arg.size() > s.size() && arg[1] != '-') {
        *out = arg.substr(s.size());
        return true;
    }
    if (longFlag[0] != 0 && arg.compare(0, l.size() + 1, l + "=") == 0) {
        *out = arg.substr(l.size() + 1);
        return true;
    }
    return false;
}

bool isFlag(const std::string &arg, const char *shortFlag, const char *longFlag)
{
    return arg == shortFlag || (longFlag[0] != 0 && arg == std::string("--") + longFlag);
}

bool parseUnsigned(const std::string &text, unsigned *out)
{
    if (text.empty() || text.find_first_not_of("0123456789") != std::string::npos) {
        return false;
    }
    *out = (unsigned) strtoul(text.c_str(), nullptr, 10);
    return true;
}

Options parse(int argc, char **argv)
{
    Options o;
    std::vector<std::string> positionals;
    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        std::string value;
        if (isFlag(arg, "-h", "help")) {
            o.help = true;
        }
        else if (arg == "--80") {
            o.eighty = true;
        }
        else if (isFlag(arg, "-v", "verbose")) {
            o.verbose = true;
        }
        else if (isFlag(arg, "-V", "version")) {
            o.version = true;
        }
        else if (isFlag(arg, "-u", "super")) {
            o.supersample = true;
        }
        else if (takeValue(argc, argv, i, "-i", "", &value)) {
            o.inputFormat = value;
        }
        else if (takeValue(argc, argv, i, "-o", "", &value)) {
            o.outputFormat = value;
        }
        else if (takeValue(argc, argv, i, "-t", "", &value)) {
            o.texture = value;
        }
        else if (takeValue(argc, argv, i, "-r", "res", &value)) {
            o.haveResolution = parseUnsigned(value, &o.resolution);
            o.bad |= !o.haveResolution;
        }
        else if (takeValue(argc, argv, i, "-s", "strat", &value)) {
            if (value == "max") {
                o.strategy = OBJ2VOXEL_MAX_STRATEGY;
            }
            else if (value == "blend") {
                o.strategy = OBJ2VOXEL_BLEND_STRATEGY;
            }
            else {
                o.bad = true;
            }
        }
        else if (takeValue(argc, argv, i, "-p", "perm", &value)) {
            o.permutation = value;
        }
        else if (takeValue(argc, argv, i, "-j", "threads", &value)) {
            o.bad |= !parseUnsigned(value, &o.threads);
        }
        else if (!arg.empty() && arg[0] == '-' && arg.size() > 1) {
            fprintf(stderr, "Flag could not be matched: %s\n", arg.c_str());
            o.bad = true;
        }
        else {
            positionals.push_back(arg);
        }
    }
    if (positionals.size() > 0) {
        o.input = positionals[0];
    }
    if (positionals.size() > 1) {
        o.output = positionals[1];
    }
    o.bad |= positionals.size() > 2;
    return o;
}

/// "xYz" -> row-major unit transform: output axis i takes input axis (c - 'x'), negated for capitals
/// (reference parsePermutation, src/main.cpp:224-262).
bool parsePermutation(const std::string &text, int out[9])
{
    if (text.size() != 3) {
        fprintf(stderr, "Invalid permutation length (%zu)\n", text.size());
        return false;
    }
    bool found[3] = {false, false, false};
    for (int i = 0; i < 3; ++i) {
        char c = text[(size_t) i];
        int sign = 1;
        if (c >= 'A' && c <= 'Z') {
            c = (char) (c - 'A' + 'a');
            sign = -1;
        }
        const int axis = c - 'x';
        if (axis < 0 || axis > 2) {
            fprintf(stderr, "Invalid permutation char: '%c'\n", c);
            return false;
        }
        found[axis] = true;
        for (int k = 0; k < 3; ++k) {
            out[i * 3 + k] = k == axis ? sign : 0;
        }
    }
    if (!(found[0] && found[1] && found[2])) {
        fprintf(stderr, "Invalid combination of permutation chars \"%s\"\n", text.c_str());
        return false;
    }
    return true;
}

}  // namespace

int main(int argc, char **argv)
{
    const auto start = std::chrono::steady_clock::now();
    const Options o = parse(argc, argv);

    if (o.version && !o.help) {
        printf("===== obj2voxel =====\nVersion:  %s\n", kVersion);
        return 0;
    }
    const bool complete = !o.bad && !o.input.empty() && !o.output.empty() && o.haveResolution;
    if (o.help || !complete) {
        printHelp(o.eighty);
        return complete ? 0 : 1;
    }
    obj2voxel_set_log_level(o.verbose ? OBJ2VOXEL_LOG_LEVEL_DEBUG : OBJ2VOXEL_LOG_LEVEL_INFO);

    int unitTransform[9];
    if (!parsePermutation(o.permutation, unitTransform)) {
        return 1;
    }

    obj2voxel_instance *instance = obj2voxel_alloc();
    std::vector<std::thread> workers;
    for (unsigned i = 0; i < o.threads; ++i) {
        workers.emplace_back(&obj2voxel_run_worker, instance);
    }
    obj2voxel_set_parallel(instance, o.threads != 0);
    obj2voxel_set_input_file(instance, o.input.c_str(), o.inputFormat.empty() ? nullptr : o.inputFormat.c_str());
    obj2voxel_set_output_file(instance, o.output.c_str(), o.outputFormat.empty() ? nullptr : o.outputFormat.c_str());

    obj2voxel_texture *texture = nullptr;
    if (!o.texture.empty()) {
        texture = obj2voxel_texture_alloc();
        if (obj2voxel_texture_load_from_file(texture, o.texture.c_str(), nullptr)) {
            obj2voxel_set_texture(instance, texture);
            printf("Loaded fallback texture \"%s\"\n", o.texture.c_str());
        }
        else {
            fprintf(stderr, "Continuing without fallback texture because it could not be loaded\n");
        }
    }
    obj2voxel_set_unit_transform(instance, unitTransform);
    obj2voxel_set_resolution(instance, o.resolution);
    obj2voxel_set_supersampling(instance, o.supersample ? 2u : 1u);
    obj2voxel_set_color_strategy(instance, o.strategy);

    const obj2voxel_error_t result = obj2voxel_voxelize(instance);

    obj2voxel_stop_workers(instance);
    for (std::thread &worker : workers) {
        worker.join();
    }
    if (texture != nullptr) {
        obj2voxel_texture_free(texture);
    }
    obj2voxel_free(instance);

    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    if (result == OBJ2VOXEL_ERR_OK) {
        printf("Done! (%.2f s)\n", seconds);
    }
    else {
        fprintf(stderr, "Failed with error %u (%.2f s)\n", (unsigned) result, seconds);
    }
    return (int) result;
}
