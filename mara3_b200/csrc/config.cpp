#include "config.hpp"
#include <iomanip>

using namespace m3b;

config_t config_t::binary_template()
{
    config_t c;
    auto& m = c.items;
    m["restart"]               = std::string();
    m["outdir"]                = std::string("data");
    m["cpi"]                   = 10.0;
    m["dfi"]                   = 1.0;
    m["tsi"]                   = 2e-3;
    m["tfinal"]                = 1.0;
    m["cfl_number"]            = 0.4;
    m["fixed_dt"]              = 0;
    m["depth"]                 = 4;
    m["begin_live_binary"]     = 1e6;
    m["conserve_linear_p"]     = 1;
    m["block_size"]            = 24;
    m["focus_factor"]          = 2.0;
    m["focus_index"]           = 2.0;
    m["threaded"]              = 1;
    m["rk_order"]              = 2;
    m["reconstruct_method"]    = std::string("plm");
    m["plm_theta"]             = 1.8;
    m["source_term_softening"] = 1.0;
    m["softening_radius"]      = 0.05;
    m["sink_radius"]           = 0.05;
    m["sink_rate"]             = 1.0;
    m["buffer_damping_rate"]   = 10.0;
    m["domain_radius"]         = 12.0;
    m["disk_radius"]           = 2.0;
    m["disk_mass"]             = 1e-3;
    m["ambient_density"]       = 1e-4;
    m["density_floor"]         = 0.0;
    m["separation"]            = 1.0;
    m["mass_ratio"]            = 1.0;
    m["eccentricity"]          = 0.0;
    m["counter_rotate"]        = 0;
    m["mach_number"]           = 10.0;
    m["axisymmetric_cs2"]      = 0;
    m["no_accretion_force"]    = 0;
    m["alpha_cutoff_radius"]   = 0.0;
    m["alpha"]                 = 0.1;
    m["nu"]                    = 0.0;
    m["mdot"]                  = 0.0;
    return c;
}

void config_t::set_value(const std::string& key, const config_value_t& value)
{
    auto it = items.find(key);
    if (it == items.end()) throw std::invalid_argument("config has no option " + key);
    if (it->second.index() != value.index()) throw std::invalid_argument("config got wrong data type for option " + key);
    it->second = value;
}

void config_t::set(const std::string& key, const std::string& value)
{
    auto it = items.find(key);
    if (it == items.end()) throw std::invalid_argument("config has no option " + key);

    switch (it->second.index())
    {
        case 0: it->second = std::stoi(value); break;   // std::stoi / std::stod throw std::invalid_argument like the reference
        case 1: it->second = std::stod(value); break;
        case 2: it->second = value; break;
    }
}

config_t config_t::from_argv(int argc, const char* const argv[])
{
    auto c = binary_template();
    auto seen = std::map<std::string, bool>();

    for (int n = 0; n < argc; ++n)
    {
        auto arg = std::string(argv[n]);
        auto eq = arg.find('=');

        if (eq == std::string::npos) continue;
        auto key = arg.substr(0, eq);
        auto val = arg.substr(eq + 1);

        if (seen.count(key)) throw std::invalid_argument("duplicate parameter " + key);
        seen[key] = true;
    }
    for (int n = 0; n < argc; ++n)      // applied in key order by the reference (std::map); order is irrelevant here
    {
        auto arg = std::string(argv[n]);
        auto eq = arg.find('=');
        if (eq != std::string::npos) c.set(arg.substr(0, eq), arg.substr(eq + 1));
    }
    return c;
}

void config_t::pretty_print(std::ostream& os, const std::string& header) const
{
    os << std::string(52, '=') << "\n";
    os << header << ":\n\n";
    std::ios orig(nullptr);
    orig.copyfmt(os);

    for (const auto& item : items)
    {
        os << '\t' << std::left << std::setw(24) << std::setfill('.') << item.first << ' ';
        std::visit([&os] (const auto& v) { os << v; }, item.second);
        os << '\n';
    }
    os << '\n';
    os.copyfmt(orig);
}
