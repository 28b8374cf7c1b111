// oat_cli.h -- the slice of lib/utility (TOMLSanitize.h, IOFormat.h) and of the two-stage
// Boost.ProgramOptions parse of the component mains that the hot-path executables need:
//   oat <component> TYPE SOURCE SINK [CONFIGURATION], `-c FILE KEY` selecting a TOML table,
//   unknown keys rejected (checkKeys, TOMLSanitize.h:101-118), CLI beats TOML (getValue :175-183),
//   arrays given as TOML strings on the command line (-H "[40,80]"), range-checked numerics (:220-277).
#pragma once
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace oat {
namespace config {

struct OptionSpec {
    std::string long_name;  // also the TOML key
    char short_name;        // 0 = none
    bool takes_value;
    std::string help;
};
using OptionTable = std::map<std::string, std::string>;  // key -> raw TOML value text

inline std::string trim(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
inline std::string unquote(const std::string &s)
{
    return (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) ? s.substr(1, s.size() - 2) : s;
}

// TOML subset: [table] headers, key = value lines (numbers, booleans, quoted strings, flat arrays), # comments
inline OptionTable getConfigTable(const std::string &file, const std::string &key)
{
    std::ifstream f(file);
    if (!f) throw std::runtime_error("Could not open configuration file '" + file + "'.");
    OptionTable t;
    std::string line, cur;
    bool found = false;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos && line.find('"') == std::string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty()) continue;
        if (line.front() == '[' && line.back() == ']' && line.find('=') == std::string::npos) {
            cur = trim(line.substr(1, line.size() - 2));
            if (cur == key) found = true;
            continue;
        }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) throw std::runtime_error("Malformed line in '" + file + "': " + line);
        if (cur == key) t[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
    if (!found) throw std::runtime_error("No configuration table named '" + key + "' was provided in the configuration file '" + file + "'");
    return t;
}
inline void checkKeys(const std::vector<OptionSpec> &options, const OptionTable &table)
{
    for (const auto &kv : table) {
        bool ok = false;
        for (const auto &o : options) ok |= (o.long_name == kv.first);
        if (!ok) throw std::runtime_error("Unknown configuration key '" + kv.first + "'.");
    }
}

struct VariableMap {
    std::map<std::string, std::string> values;  // long name -> raw text ("" for flags)
    std::vector<std::string> positional;
    bool count(const std::string &k) const { return values.count(k) != 0; }
};

// Parses argv[first..]; options may appear anywhere; "-c FILE KEY" takes two values.
inline VariableMap parse(int argc, char **argv, int first, const std::vector<OptionSpec> &options)
{
    VariableMap vm;
    for (int i = first; i < argc; ++i) {
        const std::string a = argv[i];
        const OptionSpec *spec = nullptr;
        std::string inline_val;
        bool has_inline = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string name = a.substr(2);
            const size_t eq = name.find('=');
            if (eq != std::string::npos) { inline_val = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
            for (const auto &o : options) if (o.long_name == name) spec = &o;
            if (!spec) throw std::runtime_error("unrecognised option '" + a + "'");
        } else if (a.size() == 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
            for (const auto &o : options) if (o.short_name == a[1]) spec = &o;
            if (!spec) throw std::runtime_error("unrecognised option '" + a + "'");
        } else {
            vm.positional.push_back(a);
            continue;
        }
        if (!spec->takes_value) { vm.values[spec->long_name] = ""; continue; }
        if (has_inline) { vm.values[spec->long_name] = inline_val; continue; }
        if (i + 1 >= argc) throw std::runtime_error("the required argument for option '--" + spec->long_name + "' is missing");
        vm.values[spec->long_name] = argv[++i];
        if (spec->long_name == "config") {
            if (i + 1 >= argc) throw std::runtime_error("option '--config' requires FILE and KEY");
            vm.values["config-key"] = argv[++i];
        }
    }
    return vm;
}

// CLI beats TOML (TOMLSanitize.h:175-183)
inline bool getRaw(const VariableMap &vm, const OptionTable &t, const std::string &key, std::string &out)
{
    auto i = vm.values.find(key);
    if (i != vm.values.end()) { out = i->second; return true; }
    auto j = t.find(key);
    if (j != t.end()) { out = j->second; return true; }
    return false;
}
template <typename T>
inline T to_number(const std::string &key, const std::string &raw)
{
    std::istringstream is(unquote(trim(raw)));
    T v;
    if (!(is >> v) || !(is >> std::ws).eof()) throw std::runtime_error("'" + key + "' must be a number, got '" + raw + "'.");
    return v;
}
template <typename T>
inline bool getNumericValue(const VariableMap &vm, const OptionTable &t, const std::string &key, T &value, T lower, T upper)
{
    std::string raw;
    if (!getRaw(vm, t, key, raw)) return false;
    const T v = to_number<T>(key, raw);
    if (v < lower || v > upper) {
        std::ostringstream os;
        os << "Configuration key '" << key << "' specifies a value that is out of bounds [" << lower << ", " << upper << "].";
        throw std::runtime_error(os.str());
    }
    value = v;
    return true;
}
inline bool getString(const VariableMap &vm, const OptionTable &t, const std::string &key, std::string &value)
{
    std::string raw;
    if (!getRaw(vm, t, key, raw)) return false;
    value = unquote(trim(raw));
    return true;
}
// "[a, b]" (TOML array; on the CLI it arrives as a string) with exactly n numbers
template <typename T>
inline bool getArray(const VariableMap &vm, const OptionTable &t, const std::string &key, std::vector<T> &out, size_t n)
{
    std::string raw;
    if (!getRaw(vm, t, key, raw)) return false;
    std::string s = unquote(trim(raw));
    s = trim(s);
    if (s.size() < 2 || s.front() != '[' || s.back() != ']') throw std::runtime_error("'" + key + "' must be a TOML array.");
    s = s.substr(1, s.size() - 2);
    out.clear();
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ',')) if (!trim(item).empty()) out.push_back(to_number<T>(key, item));
    if (out.size() != n) throw std::runtime_error("'" + key + "' must be a TOML array with " + std::to_string(n) + " elements.");
    return true;
}

}  // namespace config

// lib/utility/IOFormat.h:114-195 (colours only when stderr/stdout is a tty)
inline std::string whoError(const std::string &who, const std::string &msg) { return who + ": " + msg; }
inline std::string whoMessage(const std::string &who, const std::string &msg) { return who + ": " + msg; }

}  // namespace oat
