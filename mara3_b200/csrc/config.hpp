/**
 * config.hpp -- typed key=value run configuration of the `binary` subprogram.
 *
 * Mirrors the behaviour of the reference's config_t / config_template_t
 * (Mara3 src/app_config.hpp:61-186, 223-245) for the 39 keys of
 * binary::create_config_template (src/subprog_binary.cpp:57-99): every key has a
 * fixed type taken from its default (int / double / string); unknown keys,
 * duplicate command-line keys and unparsable values are errors; tokens without
 * '=' are ignored.
 */
#pragma once
#include <map>
#include <ostream>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

namespace m3b
{
    using config_value_t = std::variant<int, double, std::string>;

    class config_t
    {
    public:
        /** The `binary` template with its defaults (subprog_binary.cpp:57-99). */
        static config_t binary_template();

        /** create_run_config without restart= support (subprog_binary.cpp:155-164). */
        static config_t from_argv(int argc, const char* const argv[]);

        /** Set from a string, converting to the key's type (app_config.hpp:103-118). */
        void set(const std::string& key, const std::string& value);
        void set_value(const std::string& key, const config_value_t& value);

        int get_int(const std::string& key) const { return std::get<int>(at(key)); }
        double get_double(const std::string& key) const { return std::get<double>(at(key)); }
        const std::string& get_string(const std::string& key) const { return std::get<std::string>(at(key)); }
        bool has(const std::string& key) const { return items.count(key) > 0; }

        const std::map<std::string, config_value_t>& all() const { return items; }

        /** pretty_print (app_config.hpp:197-219): same layout as the reference. */
        void pretty_print(std::ostream& os, const std::string& header) const;

    private:
        const config_value_t& at(const std::string& key) const
        {
            auto it = items.find(key);
            if (it == items.end()) throw std::invalid_argument("config has no option " + key);
            return it->second;
        }
        std::map<std::string, config_value_t> items;   // std::map: keys print in sorted order like the reference
    };
}
